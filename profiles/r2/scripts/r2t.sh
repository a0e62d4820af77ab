#!/bin/bash
out=gpurun_out; tag=r2t; mkdir -p $out
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
b c4_f8b4 --workload c4 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/f8b3.so b c4_f8b3 --workload c4 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/f8b2.so b c4_f8b2 --workload c4 --steps 20
b c5_ub4 --workload c5 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/ub3.so b c5_ub3 --workload c5 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/ub5.so b c5_ub5 --workload c5 --steps 20
