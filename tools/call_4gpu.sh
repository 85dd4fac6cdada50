#!/bin/bash
# round 2, 4-GPU validation: all GPU tests (stepped pair matrix, stability), wire counters, pair matrix, sweep tables, bench N=4
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_4gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu_4gpu.log
timeout 300 python tools/a2a_pair_matrix.py > gpurun_out/a2a_pair_matrix.log 2>&1; tail -24 gpurun_out/a2a_pair_matrix.log
timeout 600 python tools/nvlink_counters.py > gpurun_out/nvlink_counters.log 2>&1; tail -14 gpurun_out/nvlink_counters.log
STAB_REPEATS=10 timeout 900 python tools/stability.py > gpurun_out/stability.log 2>&1; tail -6 gpurun_out/stability.log
timeout 600 python tools/sweep_tables.py > gpurun_out/sweep_tables.log 2>&1; tail -14 gpurun_out/sweep_tables.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 200 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], json.dumps(d.get('nvlink')), json.dumps(d.get('roofline_nvlink')), json.dumps(d.get('probe_round')))
PY
tail -3 gpurun_out/bench_n4.err
