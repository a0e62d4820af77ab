#!/bin/bash
out=gpurun_out; tag=r2d; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $out/${tag}_tests.log 2>&1; tail -n 6 $out/${tag}_tests.log
