#!/bin/bash
# One GPU-box pass of round 2: gpu tests, bench lines, the per-step ncu counters (bash profiles/r2_check.sh <tag> [notests])
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
if [ "$2" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $out/${tag}_tests.log 2>&1
  tail -n 15 $out/${tag}_tests.log
fi
timeout 600 python bench.py > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; echo "bench c5 rc=$?"; tail -c 600 $out/${tag}_bench_c5.err
timeout 300 python bench.py --workload c2 --steps 300 --no-cpu > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; echo "bench c2 rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file $out/${tag}_work_c5.csv --metrics $M \
    python profiles/step_capture.py --workload c5 --steps 1 > $out/${tag}_work_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 600 ncu --profile-from-start off --clock-control none --csv --log-file $out/${tag}_work_c2.csv --metrics $M \
    python profiles/step_capture.py --workload c2 --steps 2 > $out/${tag}_work_c2.log 2>&1; echo "ncu c2 rc=$?"
head -c 1500 $out/${tag}_bench_c5.json
