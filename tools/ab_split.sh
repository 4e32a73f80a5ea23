# same-box A/B: uncond / cond halves as two concurrent graph branches (CFGPP_SPLIT=1) vs one full-batch body (default)
for rep in 1 2 3; do
echo "--- SPLIT"; CFGPP_SPLIT=1 bash tools/run_diag.sh bench_unet 2>&1 | grep -E "native fused"
echo "--- NO SPLIT"; bash tools/run_diag.sh bench_unet 2>&1 | grep -E "native fused"
done
