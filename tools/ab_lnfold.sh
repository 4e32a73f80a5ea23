timeout 900 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_unet.py -m gpu -x -q 2>&1 | tail -n 6
bash tools/run_diag.sh unet_sdxl 2>&1 | grep -E "OK|BAD|EXC|std"
echo "--- NO LNFOLD"; CFGPP_NO_LNFOLD=1 bash tools/run_diag.sh bench_unet 2>&1 | grep -E "native fused"
echo "--- LNFOLD"; bash tools/run_diag.sh bench_unet 2>&1 | grep -E "native fused"
