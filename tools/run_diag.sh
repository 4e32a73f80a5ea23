#!/bin/bash
# Run each diagnostic group in its own process (a trapped kernel poisons the CUDA context) under a timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for grp in "$@"; do
  echo "##### group: $grp"
  timeout 300 python tools/diag_kernels.py $grp 2>&1 | tee gpurun_out/diag_$grp.log | tail -n 80
done
