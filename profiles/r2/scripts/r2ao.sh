#!/bin/bash
# ncu --set full of the flux stage of the last build (first operator of a step and a later one), config 5
out=gpurun_out; tag=r2ao; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled --profile-from-start off \
    -k regex:"FluxStage" -c 4 -o $out/prof_${tag}_c5 \
    python profiles/step_capture.py --workload c5 --steps 1 > $out/${tag}_ncu_full.log 2>&1; tail -n 2 $out/${tag}_ncu_full.log
ls -la $out/prof_${tag}_c5.ncu-rep
