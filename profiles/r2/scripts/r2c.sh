#!/bin/bash
out=gpurun_out; tag=r2c; mkdir -p $out
ASTREA_B200_LIB=astrea_b200/lib/variants/fma.so timeout 600 python tests/tolerance_probe.py > $out/${tag}_tolerance_fma.json 2> $out/${tag}_tolerance_fma.err
python -c "
import json; d=json.load(open('$out/${tag}_tolerance_fma.json')); print('fma worst 1 step', d['worst_one_step'], 'all', d['worst_all_steps'], 'errors', d['errors'])"
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in r.items()})
except Exception as e: print("$name failed", e)
PY
}
ASTREA_B200_LIB=astrea_b200/lib/variants/fma.so b c2_fma --workload c2 --steps 300
ASTREA_B200_LIB=astrea_b200/lib/variants/fma.so b c5_fma --workload c5 --steps 20
timeout 600 python -m pytest tests/test_ppm_dissipation.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
