#!/bin/bash
out=gpurun_out; tag=r2z; mkdir -p $out
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],4), {k:round(v,3) for k,v in r.items()}, d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
}
ASTREA_B200_LIB=astrea_b200/lib/variants/pbf.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "golden or matrix or medium" 2>&1 | tail -2
b c5 --workload c5 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/pbf.so b c5_pbf --workload c5 --steps 20
b c2 --workload c2 --steps 300
ASTREA_B200_LIB=astrea_b200/lib/variants/pbf.so b c2_pbf --workload c2 --steps 300
