#!/bin/bash
out=gpurun_out; tag=r2an; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-2} --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus ${N:-2} --steps 20 --warmup 5 --no-cpu > $out/${tag}_c5_n${N:-2}.json 2> $out/${tag}_c5_n${N:-2}.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_c5_n${N:-2}.json")); print(d["n_gpus"], d["value"], d["ms_per_step"], d["parity_check"]["ok"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
except Exception as e: print("failed", e); print(open("$out/${tag}_c5_n${N:-2}.err").read()[-2000:])
PY
