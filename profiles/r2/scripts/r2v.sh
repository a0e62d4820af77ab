#!/bin/bash
out=gpurun_out; tag=r2v; mkdir -p $out
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
b c5 --workload c5 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/pty8.so b c5_pty8 --workload c5 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/pb4.so b c5_pb4 --workload c5 --steps 20
