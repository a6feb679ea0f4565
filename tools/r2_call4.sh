#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call4.log
: > "$LOG"
timeout 300 python tools/attn_debug.py --variants 5 --batch 1 --iters 1 2>&1 | tee -a "$LOG"
timeout 300 python tools/attn_debug.py --variants 5 --batch 2 --iters 1 2>&1 | head -60 | tee -a "$LOG"
