#!/bin/bash
# two GPUs: NCCL slab parity tests (incl. Lax-Wendroff key reduction) + single-GPU LW / kernel-suite checks + bench c5 N=2
out=gpurun_out; tag=r2ad; mkdir -p $out
nvidia-smi -L | head -4
timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 3 $out/${tag}_tests.log
timeout 900 python -m pytest tests/test_gpu_kernel_suite.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "lw or Lax or unsupported or golden" > $out/${tag}_lw.log 2>&1; tail -n 3 $out/${tag}_lw.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $out/${tag}_c5.json 2> $out/${tag}_c5.err; python -c "
import json; d=json.load(open('$out/${tag}_c5.json')); print(d['value'], d['ms_per_step'], d['parity_check'], d['e2e']['value'], d['e2e']['ms_per_step'])"
