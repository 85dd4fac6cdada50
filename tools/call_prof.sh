#!/bin/bash
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'hbm_ring_kernel' -s 1 -c 4 -f -o gpurun_out/prof_hbm_modes_r01 python tools/prof_hbm.py > gpurun_out/prof_hbm.log 2>&1
tail -2 gpurun_out/prof_hbm.log
unset CUDA_VISIBLE_DEVICES
A2A_PROF_VARIANT=6 ncu --set full --clock-control none --import-source on -k regex:a2a_stagger -c 2 -f -o gpurun_out/prof_a2a_stagger_r01 python tools/prof_a2a.py > gpurun_out/prof_a2a.log 2>&1
tail -2 gpurun_out/prof_a2a.log
