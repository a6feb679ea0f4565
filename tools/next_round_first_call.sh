#!/usr/bin/env bash
# First GPU call of the next round (one gpurun, ~6 GPU-minutes):
#   here:   python tools/kernel_diag.py build            # ab/lib*.so, ~4 min of nvcc
#   then:   gpurun --timeout 900 -- bash tools/next_round_first_call.sh
# 1. parity of the experimental attention kernels (each variant in its own process: a faulting kernel poisons the CUDA
#    context of the process it ran in),
# 2. their speed against the production kernel in one process,
# 3. the differential-timing table (what each ingredient of the attention / GEMM kernels costs).
# Everything lands in gpurun_out/next_round_first_call.log.
set -u
mkdir -p gpurun_out
LOG=gpurun_out/next_round_first_call.log
: > "$LOG"
for v in 3 4; do
  echo "=== attention variant $v: operator tests" | tee -a "$LOG"
  GRL_ATTN_SPLIT=$v timeout 120 python -m pytest tests/test_gpu_tc_ops.py -q -x -k "attention" 2>&1 | tail -3 | tee -a "$LOG"
  echo "=== attention variant $v: network PSNR gates" | tee -a "$LOG"
  GRL_ATTN_SPLIT=$v timeout 200 python -m pytest tests/test_gpu_model_bf16.py -q -x 2>&1 | tail -3 | tee -a "$LOG"
done
echo "=== A/B, GRL-Base x4 SR, B=8 (variant 0 = production)" | tee -a "$LOG"
timeout 200 python tools/time_model.py --variant base --size 256 --batch 8 --precision fp16 --style init --attn-variants 0,3,4,0,3,4 2>&1 | tail -4 | tee -a "$LOG"
if [ -d ab ]; then
  echo "=== differential timing" | tee -a "$LOG"
  timeout 600 python tools/kernel_diag.py run --batch 8 2>&1 | tee -a "$LOG"
fi
