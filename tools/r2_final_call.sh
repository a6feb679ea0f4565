#!/usr/bin/env bash
# round 2: validation + measurement call (one B200): full GPU test suite, smoke, bench lines of every workload, launch
# list, ncu captures of the attention and GEMM kernels.  Results land in gpurun_out/ (copied to profiles/ by hand).
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_final.log
: > "$LOG"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tee -a "$LOG"
echo "=== pytest -m gpu" | tee -a "$LOG"
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee -a "$LOG"
echo "=== smoke" | tee -a "$LOG"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a "$LOG"
echo "=== bench cfg4 (N=1)" | tee -a "$LOG"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; tail -c 3000 gpurun_out/r2_bench_cfg4.json | tee -a "$LOG"
echo "=== bench cfg2 / cfg3 / cfg5 (N=1)" | tee -a "$LOG"
for w in cfg2 cfg3 cfg5; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; head -c 600 gpurun_out/r2_bench_$w.json | tee -a "$LOG"; echo | tee -a "$LOG"
done
echo "=== attention launches of one block (B=16) and single-tile latency, eager vs CUDA graph" | tee -a "$LOG"
timeout 300 python tools/attn_debug.py --variants 5 --batch 16 2>&1 | tail -2 | tee -a "$LOG"
timeout 300 python tools/time_model.py --variant base --size 256 --batch 1 --precision fp16 --style init --iters 5 2>&1 | tail -n 1 | tee -a "$LOG"
timeout 300 python tools/time_model.py --variant base --size 256 --batch 1 --precision fp16 --style init --iters 5 --cuda-graph 2>&1 | tail -n 1 | tee -a "$LOG"
echo "=== launch list (ncu, serialised)" | tee -a "$LOG"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --tiles-per-gpu 4 > gpurun_out/r2_ncu_bench.log 2>&1
tail -c 400 gpurun_out/r2_ncu_bench.log | tee -a "$LOG"
echo "=== ncu --set full: attn2, gemm" | tee -a "$LOG"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2 -c 3 -o gpurun_out/r2_attn2_final python tools/attn_debug.py --variants 5 --batch 8 --iters 1 > gpurun_out/r2_ncu_attn2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 8 -o gpurun_out/r2_gemm_final python tools/time_model.py --variant base --size 256 --batch 8 --precision fp16 --style init --iters 1 > gpurun_out/r2_ncu_gemm.log 2>&1
tail -n 1 gpurun_out/r2_ncu_attn2.log | tee -a "$LOG"
tail -n 1 gpurun_out/r2_ncu_gemm.log | tee -a "$LOG"
