set -u
mkdir -p gpurun_out
L=gpurun_out/r2_call29.log; : > $L
timeout 300 python tools/attn_debug.py --variants 5 --batch 2 2>&1 | tail -9 | tee -a $L
timeout 900 python tools/attn2_diag.py run --batch 16 2>&1 | tee -a $L
