bash tools/run_diag.sh gemm conv unet_tiny
echo "--- NO CLUSTER"; CFGPP_NO_CLUSTER=1 bash tools/run_diag.sh bench_unet prof_unet 2>&1 | grep -E "native|kind|total"
echo "--- CLUSTER"; bash tools/run_diag.sh bench_unet prof_unet 2>&1 | grep -E "native|kind|total|ms "
