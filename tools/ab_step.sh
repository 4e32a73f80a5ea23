# same-box A/B of the fused step over library variants:  bash tools/ab_step.sh "" ab/x/libcfgpp_b200.so ...
for rep in 1 2; do
for lib in "$@"; do
  echo "##### lib=${lib:-default} (rep $rep)"
  CFGPP_B200_LIB=$lib bash tools/run_diag.sh bench_unet 2>&1 | grep -E "native fused"
done
done
