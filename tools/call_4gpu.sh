#!/bin/bash
# round 2, 4-GPU validation: all GPU tests (stepped pair matrix, stability), pair matrix, wire counters, sweep tables, probe round, bench N=4
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_4gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu_4gpu.log
timeout 300 python tools/a2a_pair_matrix.py > gpurun_out/a2a_pair_matrix.log 2>&1; tail -24 gpurun_out/a2a_pair_matrix.log
timeout 600 python tools/nvlink_counters.py > gpurun_out/nvlink_counters.log 2>&1; tail -14 gpurun_out/nvlink_counters.log
STAB_REPEATS=10 timeout 900 python tools/stability.py > gpurun_out/stability.log 2>&1; tail -6 gpurun_out/stability.log
timeout 600 python tools/sweep_tables.py > gpurun_out/sweep_tables.log 2>&1; tail -14 gpurun_out/sweep_tables.log
timeout 300 python tools/probe_round.py 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 200 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
tail -c 2500 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
