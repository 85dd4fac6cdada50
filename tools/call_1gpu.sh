#!/bin/bash
# round 2, 1-GPU: what the driver runs at round end (tests, smoke, bench both arms)
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_1gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu_1gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['hbm_sustained']['value'], d['roofline_gemm']['achieved'], d['roofline_gemm']['classes']['uniform_philox']['ours_over_cublas'], d['probe_round'], d['clocks'])"; tail -3 gpurun_out/bench_n1.err
