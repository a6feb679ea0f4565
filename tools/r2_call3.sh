#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call3.log
: > "$LOG"
timeout 300 python tools/attn_debug.py --variants 0,5 --batch 2 2>&1 | tee -a "$LOG"
timeout 300 python tools/attn_debug.py --variants 5 --batch 16 --iters 3 2>&1 | grep -E "variant|window|identical|invariant" | tee -a "$LOG"
GRL_ATTN_SPLIT=5 timeout 300 python -m pytest tests/test_gpu_model_bf16.py -q -x -k "batch_invariance" 2>&1 | tail -3 | tee -a "$LOG"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2 -c 3 -o gpurun_out/r2_attn2_a python tools/attn_debug.py --variants 5 --batch 4 --iters 1 > gpurun_out/ncu_attn2_a.log 2>&1
tail -3 gpurun_out/ncu_attn2_a.log | tee -a "$LOG"
