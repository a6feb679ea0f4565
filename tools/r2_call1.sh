#!/usr/bin/env bash
# round 2, GPU call 1: experimental attention kernels (parity, A/B timing) + differential timing of the r1 kernel
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call1.log
: > "$LOG"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee -a "$LOG"
for v in 3 5; do
  echo "=== attention variant $v: operator tests" | tee -a "$LOG"
  GRL_ATTN_SPLIT=$v timeout 300 python -m pytest tests/test_gpu_tc_ops.py -q -x -k "attention" 2>&1 | tail -15 | tee -a "$LOG"
done
echo "=== A/B, GRL-Base x4 SR, B=8 (variant 0 = production)" | tee -a "$LOG"
timeout 300 python tools/time_model.py --variant base --size 256 --batch 8 --precision fp16 --style init --attn-variants 0,3,4,5,0,3,4,5 2>&1 | tail -8 | tee -a "$LOG"
echo "=== variant 3 network PSNR gates" | tee -a "$LOG"
GRL_ATTN_SPLIT=3 timeout 300 python -m pytest tests/test_gpu_model_bf16.py -q -x 2>&1 | tail -3 | tee -a "$LOG"
echo "=== variant 5 network PSNR gates" | tee -a "$LOG"
GRL_ATTN_SPLIT=5 timeout 300 python -m pytest tests/test_gpu_model_bf16.py -q -x 2>&1 | tail -3 | tee -a "$LOG"
if [ -d ab ]; then
  echo "=== differential timing" | tee -a "$LOG"
  timeout 600 python tools/kernel_diag.py run --batch 8 2>&1 | tee -a "$LOG"
fi
