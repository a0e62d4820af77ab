#!/bin/bash
# Evidence of one round, run on the GPU box from the repo root:  gpurun -- 'bash profiles/capture.sh r2'
# Writes into gpurun_out/: the gpu test log, bench lines of every configuration (the default line is BASELINE config 5),
# the reference arm, the ncu launch list of the default bench command, the per-step ncu counters (DRAM bytes, fp64
# instructions) of configs 5 and 2, a compute-sanitizer record and the smoke log.  profiles/summarize.py turns the ncu
# files into the tables of README.md and into traffic.json / fp64_work.json.
# Numbers printed by a run under ncu are never bench values: the bench lines come from the unprofiled runs above them.
tag=${1:-round}
out=gpurun_out
mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 3 $out/${tag}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; echo "bench c5 rc=$?"
for w in c1 c2 c3 c4; do
  steps=20; [ $w = c2 ] && steps=300; [ $w = c3 ] && steps=60; [ $w = c1 ] && steps=20000
  timeout 600 python bench.py --no-cpu --workload $w --steps $steps --warmup 5 > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err; echo "bench $w rc=$?"
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference arm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_c5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity-check > $out/${tag}_ncu_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum
for w in c5 c2 c3 c4; do
  timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file $out/${tag}_work_$w.csv --metrics $M \
      python profiles/step_capture.py --workload $w --steps 1 > $out/${tag}_work_$w.log 2>&1; echo "ncu work $w rc=$?"
done
timeout 600 compute-sanitizer --tool memcheck python tests/sanitize_smoke.py > $out/${tag}_sanitizer_memcheck.log 2>&1; tail -n 1 $out/${tag}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tests/sanitize_smoke.py > $out/${tag}_sanitizer_racecheck.log 2>&1; tail -n 1 $out/${tag}_sanitizer_racecheck.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
tail -n 3 $out/${tag}_smoke.log
