#!/bin/bash
# One GPU-box visit: bring-up selftest, GPU parity tests, bench, launch list.  Logs go to gpurun_out/.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export URNN_BENCH_TRACE=1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/gpu.txt 2>&1
echo "== selftest quick" ; timeout 120 tools/bin/tc_selftest quick > $OUT/selftest_quick.log 2>&1; RC=$?; echo "rc=$RC"; grep -E "FAIL|HANG|error|mismatch" $OUT/selftest_quick.log | grep -v "mismatches 0" | head -20
if [ $RC -ne 0 ]; then echo "selftest failed: stopping"; tail -30 $OUT/selftest_quick.log; exit 1; fi
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== bench (watchdog 150 s)"; timeout 200 python bench.py --watchdog 150 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"; cat $OUT/bench_n1.json; tail -30 $OUT/bench_n1.err
for cfg in "URNN_WIMG=0"; do
  echo "== bench value-only $cfg"
  env $cfg timeout 100 python bench.py --value-only --steps 180 --watchdog 80 > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "rc=$?"; cat $OUT/bench_$cfg.json; grep -v "^\[bench" $OUT/bench_$cfg.err | tail -12
done
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv python bench.py --value-only --steps 2 --warmup 2 > $OUT/ncu_bench.log 2>&1; echo "rc=$?"
python tools/launch_summary.py $OUT/launches.csv > $OUT/launches.txt 2>&1; tail -32 $OUT/launches.txt
