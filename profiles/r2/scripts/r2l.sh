#!/bin/bash
out=gpurun_out; tag=r2l; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_suite.py -m gpu -q --maxfail=5 -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 3 $out/${tag}_tests.log
# ncu --set full of the kernels of a step at the headline size, with source-level counters
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled --profile-from-start off \
    -k regex:"FluxStage|ReconStage|PrimBothStage|UpdateKernel" -c 8 -o $out/prof_${tag}_c5 \
    python profiles/step_capture.py --workload c5 --steps 1 > $out/${tag}_ncu_full.log 2>&1; tail -n 2 $out/${tag}_ncu_full.log
ls -la $out/prof_${tag}_c5.ncu-rep
