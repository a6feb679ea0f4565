# The last measurement call of round 2 (persistent GEMM kernels on by default).  NOT run in full: the round's GPU budget ended
# after its operator-test leg (profiles/r2_persistent_gemm_optests.log); kept as the command list for whoever re-takes the numbers.
set -u
mkdir -p gpurun_out
LOG=gpurun_out/r2_final2.log; : > "$LOG"
echo "=== pytest -m gpu (persistent GEMM kernels on by default)" | tee -a "$LOG"
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -n 4 | tee -a "$LOG"
echo "=== smoke" | tee -a "$LOG"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee -a "$LOG"
echo "=== bench cfg4" | tee -a "$LOG"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; head -c 400 gpurun_out/r2_bench_cfg4.json | tee -a "$LOG"; echo | tee -a "$LOG"
for w in cfg2 cfg3 cfg5; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; head -c 300 gpurun_out/r2_bench_$w.json | tee -a "$LOG"; echo | tee -a "$LOG"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcp -s 40 -c 8 -o gpurun_out/r2_gemm_final python tools/time_model.py --variant base --size 256 --batch 8 --precision fp16 --style init --iters 1 > gpurun_out/r2_ncu_gemm.log 2>&1
tail -n 1 gpurun_out/r2_ncu_gemm.log | tee -a "$LOG"
