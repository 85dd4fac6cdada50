#!/bin/bash
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_elapsed.avg,dram__cycles_active.avg,dram__cycles_active_read.avg,dram__cycles_active_write.avg,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:hbm_ring_kernel -c 40 --csv --log-file gpurun_out/ncu_hbm_dram_counters.csv python tools/prof_hbm.py > gpurun_out/prof_hbm.log 2>&1; tail -1 gpurun_out/prof_hbm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hbm_ring_kernel -s 12 -c 2 -f -o gpurun_out/prof_hbm_copy_r02 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gemm --no-probe-round > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
