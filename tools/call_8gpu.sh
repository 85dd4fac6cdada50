#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep_tables.py 2>&1 | tail -32
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_8.txt
