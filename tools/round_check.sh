#!/bin/bash
# What the driver runs at round end (GPU tests, smoke, bench) + the ncu evidence, in one gpurun call.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "##### pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 25 | tee gpurun_out/pytest_gpu.log
echo "##### smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 5
echo "##### bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
echo "##### ncu launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 1 > gpurun_out/ncu_list.log 2>&1; tail -n 2 gpurun_out/ncu_list.log
echo "##### ncu full: gemm"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_kernel -s 200 -c 12 -o gpurun_out/prof_gemm -f python tools/ncu_target.py 1 > gpurun_out/ncu_gemm.log 2>&1; tail -n 2 gpurun_out/ncu_gemm.log
echo "##### ncu full: attention"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_kernel -s 30 -c 4 -o gpurun_out/prof_attn -f python tools/ncu_target.py 1 > gpurun_out/ncu_attn.log 2>&1; tail -n 2 gpurun_out/ncu_attn.log
fi
ls -la gpurun_out | head -30
