#!/bin/bash
# one-GPU check: parity suite, GEMM variants, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_1.txt
for v in 1 2; do B200PROBE_GEMM_VARIANT=$v python tools/gemm_bench.py 2>&1 | tail -8; done | tee gpurun_out/gemm_bench.txt
python bench.py --steps 500 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
