#!/bin/bash
out=gpurun_out; tag=r2e; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_suite.py -m gpu -q -x -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 4 $out/${tag}_tests.log
timeout 300 compute-sanitizer --tool memcheck python tests/sanitize_smoke.py > $out/${tag}_sanitizer.log 2>&1; tail -n 3 $out/${tag}_sanitizer.log
Q="--no-cpu --no-e2e --no-parity-check"
b() { name=$1; shift; timeout 300 python bench.py $Q "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); r=d["roofline"]["class_ms_per_step"]
    print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in r.items()})
except Exception as e: print("$name failed", e)
PY
}
b c2_bulk --workload c2 --steps 300
b c2_nobulk --workload c2 --steps 300 --no-recon-bulk
b c5_bulk --workload c5 --steps 20
b c5_nobulk --workload c5 --steps 20 --no-recon-bulk
for v in rb4 rs2 rs4; do
ASTREA_B200_LIB=astrea_b200/lib/variants/$v.so b c5_$v --workload c5 --steps 20
ASTREA_B200_LIB=astrea_b200/lib/variants/$v.so b c2_$v --workload c2 --steps 300
done
b c3_bulk --workload c3 --steps 40
b c3_nobulk --workload c3 --steps 40 --no-recon-bulk
b c4_bulk --workload c4 --steps 20
b c4_nobulk --workload c4 --steps 20 --no-recon-bulk
