#!/bin/bash
# GEMM skew experiment on the experimental build of the library
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
export B200PROBE_LIB=$PWD/k3s-nvidia_b200/libb200probe_exp.so
for sk in 0 2 1 3; do echo "== skew $sk"; B200PROBE_GEMM_SKEW=$sk timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "pair512 or headline" 2>&1 | tail -3; done
GEMM_TUNE_QUICK=1 timeout 900 python tools/gemm_tune.py 2>&1 | tail -26
