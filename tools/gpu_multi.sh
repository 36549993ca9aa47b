#!/bin/bash
# Multi-GPU evidence run (gpurun --gpus N): band-sharding parity at this world size, weak and strong scaling bench lines.
# usage: tools/gpu_multi.sh N [tests]
set -u
N=${1:-8}; OUT=gpurun_out; mkdir -p $OUT
if [ "${2:-}" = "tests" ]; then
  timeout 400 python -m pytest tests/test_gpu_multi.py -x -q > $OUT/r2_pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?"; tail -3 $OUT/r2_pytest_multi_n$N.log
fi
run() {  # name, extra args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) \
      bench.py --gpus $N $2 > $OUT/r2_bench_n${N}_$1.json 2> $OUT/r2_bench_n${N}_$1.err
  echo "bench $1 rc=$?"; cat $OUT/r2_bench_n${N}_$1.json | cut -c1-700
}
run weak ""
run strong "--scaling strong --height 4096 --width 4096 --steps 20 --warmup 3"
