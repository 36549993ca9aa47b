#!/bin/bash
# Final single-GPU evidence run: tests, bench, launch list, ncu --set full of the roofline kernel (cell step).
set -u
mkdir -p gpurun_out; OUT=gpurun_out
export URNN_BENCH_TRACE=1
timeout 120 tools/bin/tc_selftest quick > $OUT/selftest_quick.log 2>&1; echo "selftest rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python bench.py --watchdog 250 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; cat $OUT/bench_n1.json
timeout 300 python bench.py --math fp32 --steps 30 --no-cpu-baseline --watchdog 250 > $OUT/bench_n1_fp32.json 2>/dev/null; echo "bench fp32 rc=$?"; cat $OUT/bench_n1_fp32.json
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_ref.json 2>/dev/null; echo "ref rc=$?"; cat $OUT/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file $OUT/launches.csv python bench.py --value-only --steps 2 --warmup 2 > $OUT/ncu_bench.log 2>&1
python tools/launch_summary.py $OUT/launches.csv > $OUT/launches.txt 2>&1; tail -3 $OUT/launches.txt
timeout 300 ncu --set full --import-source on --clock-control none -k "regex:gemm_gn_kernel|cgru_blend" -s 6 -c 3 -f -o $OUT/cell python tools/prof_cell.py > $OUT/ncu_cell.log 2>&1; echo "ncu cell rc=$?"
ncu -i $OUT/cell.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_traffic.py > $OUT/traffic.json; cat $OUT/traffic.json | head -40
