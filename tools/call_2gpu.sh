#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_a2a.py tests/test_gpu_hbm.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/sweep_tables.py 2>&1 | tail -28
