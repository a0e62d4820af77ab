#!/bin/bash
out=gpurun_out; tag=r2; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "golden or odd or constrained" 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; echo "bench c5 rc=$?"
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_c5.json")); r=d["roofline"]; print(d["value"], d["ms_per_step"], r["frac"], r["fp64"]["pipe_busy"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r["class_ms_per_step"])
PY
