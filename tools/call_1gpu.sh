#!/bin/bash
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -5
GEMM_TUNE_QUICK=1 timeout 900 python tools/gemm_tune.py 2>&1 | tail -26
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm|nvjet' -s 2 -c 2 -f -o gpurun_out/prof_gemm_r02e python tools/prof_gemm.py > gpurun_out/prof_gemm.log 2>&1; tail -2 gpurun_out/prof_gemm.log
