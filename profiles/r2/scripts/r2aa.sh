#!/bin/bash
out=gpurun_out; tag=r2aa; mkdir -p $out
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
b c5_bulk --workload c5 --steps 20
b c5_plain --workload c5 --steps 20 --no-prim-bulk
b c2_bulk --workload c2 --steps 50
b c2_plain --workload c2 --steps 50 --no-prim-bulk
b c4_bulk --workload c4 --steps 20
b c4_plain --workload c4 --steps 20 --no-prim-bulk
b c3_bulk --workload c3 --steps 20
b c3_plain --workload c3 --steps 20 --no-prim-bulk
