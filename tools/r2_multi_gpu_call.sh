#!/usr/bin/env bash
# round 2: multi-GPU bench lines (gpurun --gpus N -- 'bash tools/r2_multi_gpu_call.sh N').  One process per GPU (torchrun),
# NCCL; cfg4 = batch-sharded tiles (weak, and strong with the same global batch as N=1), cfg5 = one frame, tiles sharded.
set -u
N=${1:-2}
mkdir -p gpurun_out
LOG=gpurun_out/r2_multi_n$N.log
: > "$LOG"
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv | tee -a "$LOG"
run() {  # name, extra bench args
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus "$N" --no-cpu-baseline --no-extras "$@" > gpurun_out/r2_bench_${name}_n$N.json 2> gpurun_out/r2_bench_${name}_n$N.err
  echo "--- $name rc=$?" | tee -a "$LOG"
  tail -n 1 gpurun_out/r2_bench_${name}_n$N.json | head -c 1500 | tee -a "$LOG"; echo | tee -a "$LOG"
  tail -n 3 gpurun_out/r2_bench_${name}_n$N.err | tee -a "$LOG"
}
if [ "$N" -le 2 ]; then  # single-GPU checks that ride along on the small box
  timeout 600 python -m pytest tests/test_gpu_tc_ops.py tests/test_gpu_native_shapes.py -q -k "attention or lazy_rescale or cfg3" 2>&1 | tail -n 2 | tee -a "$LOG"
  timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err
  head -c 300 gpurun_out/r2_bench_cfg3.json | tee -a "$LOG"; echo | tee -a "$LOG"
fi
run cfg5 --workload cfg5 --steps 5 --warmup 3
if [ "$N" -le 2 ]; then
  run cfg4_weak --workload cfg4 --steps 5 --warmup 3
  run cfg4_strong --workload cfg4 --scaling strong --steps 5 --warmup 3
fi
