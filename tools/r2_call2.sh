#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call2.log
: > "$LOG"
echo "=== attention variant 5: operator tests" | tee -a "$LOG"
GRL_ATTN_SPLIT=5 timeout 300 python -m pytest tests/test_gpu_tc_ops.py -q -k "attention" 2>&1 | tail -25 | tee -a "$LOG"
echo "=== A/B, GRL-Base x4 SR, B=8" | tee -a "$LOG"
timeout 300 python tools/time_model.py --variant base --size 256 --batch 8 --precision fp16 --style init --attn-variants 0,3,0,3 2>&1 | tail -4 | tee -a "$LOG"
timeout 300 python tools/time_model.py --variant base --size 256 --batch 8 --precision fp16 --style init --attn-variants 0,5,0,5 2>&1 | tail -4 | tee -a "$LOG"
echo "=== variant 5 network PSNR gates" | tee -a "$LOG"
GRL_ATTN_SPLIT=5 timeout 300 python -m pytest tests/test_gpu_model_bf16.py -q -x 2>&1 | tail -3 | tee -a "$LOG"
echo "=== native shapes (fp32 + tc)" | tee -a "$LOG"
timeout 600 python -m pytest tests/test_gpu_native_shapes.py -q -s -k "cfg4_init or cfg3_init" 2>&1 | grep -E "cfg|passed|failed|Error" | tee -a "$LOG"
