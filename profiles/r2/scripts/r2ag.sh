#!/bin/bash
# host topology of the box and the effect of binding each rank to its GPU's socket on the end-to-end rate
out=gpurun_out; tag=r2ag; n=${1:-2}; mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1; cat $out/${tag}_topo.txt | head -20
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" ; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nproc
for mode in bind nobind; do
  flag=""; [ $mode = nobind ] && flag="--no-host-bind"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu $flag > $out/${tag}_${mode}_n$n.json 2> $out/${tag}_${mode}_n$n.err; echo "$mode rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_${mode}_n$n.json")); print("$mode", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["host_binding"])
except Exception as e: print("failed", e); print(open("$out/${tag}_${mode}_n$n.err").read()[-2000:])
PY
done
