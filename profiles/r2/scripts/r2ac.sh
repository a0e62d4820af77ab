#!/bin/bash
mkdir -p gpurun_out
for cfg in "4096 8" "1024 8" "2048 8" "1024 12" "1024 16" "512 8" "2048 12"; do set -- $cfg
for w in c2 c5; do ASTREA_PIECE_KB=$1 ASTREA_LANES=$2 timeout 600 python profiles/e2e_breakdown.py $w 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 KB x $2 lanes', d['workload'], 'up', round(d['upload_pageable_ms'],2), 'down', round(d['download_pageable_ms'],2), 'pinned', round(d['upload_pinned_ms'],2), round(d['download_pinned_ms'],2))"
done; done
