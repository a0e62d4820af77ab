#!/bin/bash
out=gpurun_out; tag=r2g; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_suite.py -m gpu -q --maxfail=5 -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 4 $out/${tag}_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_kernel_suite.py::test_run_steps_persistent_replay" -m gpu -q -x -p no:cacheprovider > $out/${tag}_sanitizer_replay.log 2>&1; grep -m 12 -A12 "Invalid\|ERROR SUMMARY" $out/${tag}_sanitizer_replay.log | head -60
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
b c4 --workload c4 --steps 20
b c3 --workload c3 --steps 40
