#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_call8.log
: > "$LOG"
timeout 900 python tools/attn2_diag.py run --batch 16 2>&1 | tee -a "$LOG"
GRL_ATTN2_NO_SMEM_BIAS=1 timeout 300 python tools/attn_debug.py --variants 5 --batch 16 --iters 3 2>&1 | grep -E "window [0-9]" | sed 's/^/no-smem-bias: /' | tee -a "$LOG"
timeout 300 python -m pytest tests/test_gpu_metrics.py -q 2>&1 | tail -3 | tee -a "$LOG"
