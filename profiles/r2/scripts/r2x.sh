#!/bin/bash
out=gpurun_out; tag=r2x; mkdir -p $out
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
b c5_bulk --workload c5 --steps 20
b c5_nobulk --workload c5 --steps 20 --no-recon-bulk
b c2_bulk --workload c2 --steps 300
b c2_nobulk --workload c2 --steps 300 --no-recon-bulk
b c4_bulk --workload c4 --steps 20
b c4_nobulk --workload c4 --steps 20 --no-recon-bulk
