#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call5.log
: > "$LOG"
GRL_ATTN_SPLIT=5 timeout 300 python -m pytest tests/test_gpu_tc_ops.py -q -k "attention" 2>&1 | tail -5 | tee -a "$LOG"
timeout 300 python tools/attn_debug.py --variants 5 --batch 2 --iters 3 2>&1 | head -40 | tee -a "$LOG"
timeout 300 python tools/attn_debug.py --variants 0,5 --batch 16 --iters 3 2>&1 | grep -v "    b " | tee -a "$LOG"
GRL_ATTN_SPLIT=5 timeout 300 python -m pytest tests/test_gpu_model_bf16.py -q -x 2>&1 | tail -3 | tee -a "$LOG"
