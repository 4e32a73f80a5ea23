for cap in 0 74 112; do echo "--- SM_CAP=$cap"; CFGPP_SM_CAP=$cap bash tools/run_diag.sh bench_unet 2>&1 | grep -E "native fused"; done
