#!/bin/bash
# ncu evidence of ONE fused CFG++ step (SDXL 1024x1024, batch 2 => UNet batch 4) in one gpurun call (1 GPU):
#   launches.csv   every launch with its device time (shares)
#   metrics.csv    every launch with DRAM bytes, tensor / XU pipe activity, issue activity (per-kernel-class table)
#   prof_*.ncu-rep `--set full` captures of a few launches of every kernel class
# Summarise here with: python tools/summarize_profiles.py r02 "<note>"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python tools/ncu_target.py 1"
COMMON="--profile-from-start off --clock-control none"
echo "##### launch list"
timeout 600 ncu $COMMON --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv $T > gpurun_out/ncu_list.log 2>&1; tail -n 1 gpurun_out/ncu_list.log
echo "##### per-launch metrics"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
timeout 900 ncu $COMMON --metrics $M --csv --log-file gpurun_out/metrics.csv $T > gpurun_out/ncu_metrics.log 2>&1; tail -n 1 gpurun_out/ncu_metrics.log
full() {  # name regex skip count
  timeout 600 ncu $COMMON --set full --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1 -f $T > gpurun_out/ncu_$1.log 2>&1
  echo "full $1: $(tail -n 1 gpurun_out/ncu_$1.log)"
}
echo "##### ncu --set full"
full conv gemm_kernel 0 2          # down_blocks.0.resnets.0 conv1 / conv2 (implicit-GEMM conv3x3 at 128x128x320)
full gemm gemm_kernel 200 12       # one 1280-channel transformer block (qkv, to_out, to_q, to_out, geglu, ff.out) x2
full attn attn_kernel 30 2         # self-attention, N = 1024 x 20 heads
full xattn xattn_kernel 30 2       # cross-attention, 77 keys
full gn gn_ 0 4                    # gn_stats + gn_apply at 128x128x320 (x2)
full convio conv_ 0 2              # conv_in_kernel, conv_out_step_kernel (the fused CFG++ / DDIM epilogue)
full small "small_linear|sincos|select_step|upsample" 0 8
ls -la gpurun_out | head -40
