#!/bin/bash
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
tools/build/smid_probe > gpurun_out/smid_probe.txt 2>&1; head -3 gpurun_out/smid_probe.txt; grep -c "not one TPC" gpurun_out/smid_probe.txt; awk 'NR>2{print $2}' gpurun_out/smid_probe.txt | tr '\n' ' ' | head -c 600; echo
export B200PROBE_LIB=$PWD/k3s-nvidia_b200/libb200probe_exp.so
B200PROBE_GEMM_SMID_MAP=1 timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "pair512-stg or headline" 2>&1 | tail -3
GEMM_TUNE_QUICK=1 timeout 900 python tools/gemm_tune.py 2>&1 | tail -24
