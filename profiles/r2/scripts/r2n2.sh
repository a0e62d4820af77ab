#!/bin/bash
# two-GPU pass: NCCL slab parity tests and the bench line with its in-run parity check
out=gpurun_out; tag=${1:-r2n2}; n=${2:-2}; mkdir -p $out
nvidia-smi -L | head -8
if [ "$n" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 3 $out/${tag}_tests.log; fi
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); print("$name", d["value"], d["ms_per_step"], d["parity_check"], d["e2e"]["value"] if d.get("e2e") else None)
except Exception as e: print("$name failed", e); print(open("$out/${tag}_$name.err").read()[-1500:])
PY
}
run c5 --steps 20 --warmup 5
run c2 --workload c2 --steps 300 --warmup 20 --no-e2e
