#!/bin/bash
out=gpurun_out; tag=r2am; mkdir -p $out
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
b c5_fb5 --workload c5 --steps 20
b c2_fb5 --workload c2 --steps 100
export ASTREA_B200_LIB=astrea_b200/lib/variants/fb6.so
b c5_fb6 --workload c5 --steps 20
b c2_fb6 --workload c2 --steps 100
b c3_fb6 --workload c3 --steps 20
