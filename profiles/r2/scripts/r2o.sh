#!/bin/bash
out=gpurun_out; tag=r2o; mkdir -p $out
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
b c5_t96 --workload c5 --steps 20 --threads-2d 96
b c5_t128 --workload c5 --steps 20
b c5_seg32 --workload c5 --steps 20 --segment-2d 32
b c5_seg128 --workload c5 --steps 20 --segment-2d 128
b c5_seg256 --workload c5 --steps 20 --segment-2d 256
