#!/bin/bash
out=gpurun_out; tag=r2ap; mkdir -p $out
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 200 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
ASTREA_B200_LIB=astrea_b200/lib/variants/pf16.so b c5_pf16 --workload c5 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/pf32.so b c5_pf32 --workload c5 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/pf16.so b c2_pf16 --workload c2 --steps 100
