#!/bin/bash
# end-of-round refresh on one GPU: the whole gpu test suite, the bench lines whose host seam changed, smoke
out=gpurun_out; tag=r2ae; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 3 $out/${tag}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; echo "bench c5 rc=$?"
timeout 600 python bench.py --no-cpu --workload c2 --steps 300 --warmup 5 > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; echo "bench c2 rc=$?"
python - <<PY
import json
for w in ("c5","c2"):
    d=json.load(open("$out/${tag}_bench_%s.json"%w)); print(w, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"])
PY
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; tail -n 2 $out/${tag}_smoke.log
