#!/bin/bash
mkdir -p gpurun_out
for w in c2 c5; do timeout 600 python profiles/e2e_breakdown.py $w 2> gpurun_out/r2ab_$w.err | tee gpurun_out/r2ab_$w.json; tail -3 gpurun_out/r2ab_$w.err; done
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --workload c2 --no-cpu --steps 20 > gpurun_out/r2ab_bench_c2.json 2> gpurun_out/r2ab_bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/r2ab_bench_c2.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'])"
