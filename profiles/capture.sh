#!/bin/bash
# Evidence of one round, run on the GPU box from the repo root:  gpurun -- 'bash profiles/capture.sh r1s'
# Writes bench lines of every configuration, the reference arm, the ncu launch list and one `--set full` capture of the
# four kernels of a step into gpurun_out/; profiles/summarize.py turns the ncu files into the tables of README.md.
# Numbers printed by a run under ncu are never bench values: the bench lines come from the unprofiled runs above them.
tag=${1:-round}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
for w in c1 c3 c4 c5; do python bench.py --no-cpu --workload $w > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err; done
python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"FluxStage|ReconStage|PrimBothStage|UpdateKernel" -c 6 -o $out/prof_${tag} \
    python bench.py --steps 1 --warmup 1 --no-cpu > $out/${tag}_ncu_full.log 2>&1
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
tail -n 3 $out/${tag}_smoke.log
