#!/bin/bash
out=gpurun_out; tag=r2w; mkdir -p $out
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
ASTREA_B200_LIB=astrea_b200/lib/variants/rbb3.so b c5_rbb3 --workload c5 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/rbb3.so b c4_rbb3 --workload c4 --steps 20
b c3 --workload c3 --steps 40
ASTREA_B200_LIB=astrea_b200/lib/variants/rw4.so b c3_rw4 --workload c3 --steps 40
ASTREA_B200_LIB=astrea_b200/lib/variants/rw6.so b c3_rw6 --workload c3 --steps 40
