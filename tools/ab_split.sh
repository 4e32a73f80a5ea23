timeout 900 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_unet.py -m gpu -x -q 2>&1 | tail -n 6
echo "--- NO SPLIT"; CFGPP_NO_SPLIT=1 bash tools/run_diag.sh bench_unet 2>&1 | grep -E "native"
echo "--- SPLIT"; bash tools/run_diag.sh bench_unet 2>&1 | grep -E "native"
