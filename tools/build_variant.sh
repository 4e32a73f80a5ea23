#!/bin/bash
# Build a variant of libcfgpp_b200.so with extra -D flags into ab/<name>/libcfgpp_b200.so (git-ignored, travels with
# gpurun) for same-box A/B runs:  tools/build_variant.sh lo_issuer -DCFGPP_GEMM_HI_ISSUER=0 -DCFGPP_ATTN_HI_ISSUER=0
# Use it with  CFGPP_B200_LIB=ab/<name>/libcfgpp_b200.so python tools/...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=ab/$name
mkdir -p $out/obj
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
pids=()
for src in cfgpp_b200/csrc/*.cu; do
  o=$out/obj/$(basename ${src%.cu}).o
  $NVCC -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler \
    -fvisibility=hidden --expt-relaxed-constexpr -I include -I cfgpp_b200/csrc "$@" -c $src -o $o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $out/libcfgpp_b200.so $out/obj/*.o
echo $out/libcfgpp_b200.so
