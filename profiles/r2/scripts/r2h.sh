#!/bin/bash
out=gpurun_out; tag=r2h; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernel_suite.py -m gpu -q --maxfail=5 -p no:cacheprovider -k "run_steps or async_stepping or integrator" > $out/${tag}_tests.log 2>&1; tail -n 4 $out/${tag}_tests.log
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
b c1 --workload c1 --steps 20000
b c1_8k --workload c1 --cells 8192 --steps 5000
