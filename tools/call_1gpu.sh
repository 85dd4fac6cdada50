#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_1.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 500 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; s=open('gpurun_out/bench_n1.json').read().strip().splitlines(); print(len(s),'line(s)'); d=json.loads(s[-1]); print(d['value'], d['roofline']['frac'], d['e2e'], d['e2e_hostbuf']['value'], d['hbm_read_gbs'], d['hbm_write_gbs'])"; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | head -c 400
