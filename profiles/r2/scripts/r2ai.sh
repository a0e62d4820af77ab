#!/bin/bash
# bench lines once more, now that fp64_work.json / traffic.json hold the counters of this build
out=gpurun_out; tag=r2; mkdir -p $out
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; echo "bench c5 rc=$?"
for w in c2 c3 c4; do
  steps=20; [ $w = c2 ] && steps=300; [ $w = c3 ] && steps=60
  timeout 600 python bench.py --no-cpu --workload $w --steps $steps --warmup 5 > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err; echo "bench $w rc=$?"
done
python - <<PY
import json
for w in ("c5","c2","c3","c4"):
    d=json.load(open("$out/${tag}_bench_%s.json"%w)); r=d["roofline"]; print(w, d["value"], d["ms_per_step"], r["frac"], r["fp64"]["pipe_busy"], r["fp64"]["fp64_warp_inst_per_cell_update"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
