#!/bin/bash
# eight GPUs: the default bench line under torchrun, as the driver launches it
out=gpurun_out; tag=r2af; n=${1:-8}; mkdir -p $out
nvidia-smi -L | wc -l; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_c5_n$n.json 2> $out/${tag}_c5_n$n.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_c5_n$n.json")); print(d["n_gpus"], d["value"], d["ms_per_step"], d["parity_check"]["ok"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"])
except Exception as e: print("failed", e); print(open("$out/${tag}_c5_n$n.err").read()[-2000:])
PY
