#!/bin/bash
out=gpurun_out; tag=r2q; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_suite.py tests/test_gpu_fullsize.py -m gpu -q --maxfail=5 -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 2 $out/${tag}_tests.log
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
b c4 --workload c4 --steps 20
