timeout 600 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_unet.py -m gpu -x -q 2>&1 | tail -n 6
echo "--- NO PDL"; CFGPP_NO_PDL=1 timeout 300 python -m pytest tests/test_gpu_samplers.py -m gpu -x -q -k "fused_graph" 2>&1 | tail -n 3
CFGPP_NO_PDL=1 bash tools/run_diag.sh bench_unet
echo "--- PDL"; bash tools/run_diag.sh bench_unet
