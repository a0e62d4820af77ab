#!/bin/bash
out=gpurun_out; tag=r2b; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_suite.py tests/test_gpu_fullsize.py -m gpu -q --maxfail=10 -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 4 $out/${tag}_tests.log
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in r.items()})
except Exception as e: print("$name failed", e)
PY
}
b c2_bt128 --workload c2 --steps 300
b c2_bt64 --workload c2 --steps 300 --threads-2d 64
b c2_bt96 --workload c2 --steps 300 --threads-2d 96
ASTREA_B200_LIB=astrea_b200/lib/variants/nobt.so b c2_nobt --workload c2 --steps 300
b c5_bt128 --workload c5 --steps 20
b c5_bt64 --workload c5 --steps 20 --threads-2d 64
ASTREA_B200_LIB=astrea_b200/lib/variants/nobt.so b c5_nobt --workload c5 --steps 20
b c3 --workload c3 --steps 40
b c4 --workload c4 --steps 20
b c1 --workload c1 --steps 2000
