#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cond" 2>&1 | tail -4
timeout 600 python tools/cond_train_bench.py 8 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print(d['cond_train_step'])"
timeout 300 python tools/cond_bench.py 8 2>&1 | head -2
