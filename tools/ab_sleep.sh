bash tools/run_diag.sh gemm attn unet_tiny bench_unet 2>&1 | grep -E "===|native fused|BAD|EXC"
timeout 400 python tools/power_probe.py 2>&1 | grep -E "matmul|ours|gemm|attention"
