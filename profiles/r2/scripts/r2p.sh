#!/bin/bash
out=gpurun_out; tag=r2p; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 -p no:cacheprovider -k "drop_in or golden_vs or full_size" > $out/${tag}_tests.log 2>&1; tail -n 2 $out/${tag}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; echo rc=$?
timeout 600 python bench.py --workload c2 --steps 300 --warmup 20 --no-cpu > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; echo rc=$?
python - <<PY
import json
for w in ("c5","c2"):
    d=json.load(open("$out/${tag}_bench_%s.json"%w)); print(w, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["fp64"])
PY
