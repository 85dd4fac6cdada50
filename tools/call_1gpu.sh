#!/bin/bash
# round 2, 1-GPU: what the driver runs (tests, smoke, bench) + the ncu evidence for the bench line
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_1gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu_1gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r02_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_under_ncu.log 2>&1; tail -2 gpurun_out/smoke_under_ncu.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_elapsed.avg,dram__cycles_active.avg,dram__cycles_active_read.avg,dram__cycles_active_write.avg,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:'hbm_ring_kernel<4>|hbm_ring_kernel<1>|hbm_ring_kernel<2>' -c 40 --csv --log-file gpurun_out/ncu_hbm_dram_counters.csv python tools/prof_hbm.py > gpurun_out/prof_hbm.log 2>&1; tail -1 gpurun_out/prof_hbm.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_bench_n1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gemm --no-probe-round > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hbm_ring_kernel<4>' -s 5 -c 2 -f -o gpurun_out/prof_hbm_copy_r02 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gemm --no-probe-round > /dev/null 2>&1
( time timeout 900 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json; tail -4 gpurun_out/bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err; cat gpurun_out/bench_ref.json | cut -c1-1200
