# same-box A/B over library variants built with tools/build_variant.sh:  bash tools/ab_libs.sh "" ab/x/libcfgpp_b200.so ...
for lib in "$@"; do
  echo "##### lib=${lib:-default}"
  CFGPP_B200_LIB=$lib timeout 200 python tools/gemm_timeline.py 2 2>&1 | grep -E "GEMM|first_full|tile0_lastkb|lasttile|epi|exit"
  CFGPP_B200_LIB=$lib timeout 400 python tools/power_probe.py 2>&1 | grep -E "ours|attention|gemm"
  CFGPP_B200_LIB=$lib bash tools/run_diag.sh bench_unet 2>&1 | grep -E "native fused"
done
