#!/bin/bash
# round 2, 1-GPU validation + HBM evidence: GPU tests, policy/map/phase experiments on the copy kernel, DRAM counters per mode
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_1gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu_1gpu.log
timeout 600 python tools/hbm_evidence.py > gpurun_out/hbm_evidence.log 2>&1; tail -3 gpurun_out/hbm_evidence.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_elapsed.avg,dram__cycles_active.avg,dram__cycles_active_read.avg,dram__cycles_active_write.avg,dram__throughput.avg.pct_of_peak_sustained_elapsed,fbpa__dram_read_throughput.avg.pct_of_peak_sustained_elapsed,fbpa__dram_write_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct,dram__cycles_elapsed.avg.per_second
timeout 600 ncu --metrics $M --clock-control none -k regex:'hbm_ring_kernel' -c 4 --csv --log-file gpurun_out/ncu_hbm_dram_counters.csv python tools/prof_hbm.py > gpurun_out/prof_hbm.log 2>&1
tail -2 gpurun_out/prof_hbm.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv
