#!/bin/bash
# One GPU-box visit: bring-up selftest, GPU parity tests, bench, launch list.  Logs go to gpurun_out/.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/gpu.txt 2>&1
echo "== selftest quick" ; timeout 120 tools/bin/tc_selftest quick > $OUT/selftest_quick.log 2>&1; RC=$?; echo "rc=$RC"; grep -E "FAIL|HANG|error|mismatch" $OUT/selftest_quick.log | head -20
if [ $RC -ne 0 ]; then
  echo "selftest failed -> running the rest with URNN_BULK=0"; export URNN_BULK=0
else
  echo "== selftest full"; timeout 300 tools/bin/tc_selftest > $OUT/selftest_full.log 2>&1; echo "rc=$?"; grep -E "^==|time |FAIL|HANG" $OUT/selftest_full.log | tail -40
fi
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"; cat $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
echo "== bench simt/noreverse (value only)"
URNN_BULK=0 timeout 300 python bench.py --value-only --steps 60 > $OUT/bench_simt.json 2>/dev/null; cat $OUT/bench_simt.json
URNN_REVERSE=0 timeout 300 python bench.py --value-only --steps 60 > $OUT/bench_norev.json 2>/dev/null; cat $OUT/bench_norev.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv python bench.py --value-only --steps 2 --warmup 2 > $OUT/ncu_bench.log 2>&1; echo "rc=$?"
python tools/launch_summary.py $OUT/launches.csv > $OUT/launches.txt 2>&1; tail -40 $OUT/launches.txt
