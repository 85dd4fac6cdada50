#!/bin/bash
# round 2, 8-GPU validation: all GPU tests, pair matrix, wire counters, NVLink stability, sweep tables, native probe round, bench N=8
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_8gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu_8gpu.log
timeout 300 python tools/a2a_pair_matrix.py > gpurun_out/a2a_pair_matrix.log 2>&1; tail -42 gpurun_out/a2a_pair_matrix.log
timeout 600 python tools/nvlink_counters.py > gpurun_out/nvlink_counters.log 2>&1; tail -12 gpurun_out/nvlink_counters.log
STAB_ONLY_NVLINK=1 timeout 600 python tools/stability.py > gpurun_out/stability.log 2>&1; tail -5 gpurun_out/stability.log
timeout 600 python tools/sweep_tables.py > gpurun_out/sweep_tables.log 2>&1; tail -14 gpurun_out/sweep_tables.log
B200PROBE_IGNORE_TENANTS=1 timeout 300 host/cpp/build/b200-device-plugin --probe-once --features-dir gpurun_out/features_g8 > gpurun_out/probe_once_g8.txt 2> gpurun_out/probe_once_g8.err; grep -c . gpurun_out/probe_once_g8.txt; grep -E "b200probe\.(healthy|nvlink-|hbm-healthy|gemm-healthy)" gpurun_out/probe_once_g8.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 200 --warmup 5 ) > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['hbm_sustained'], json.dumps(d.get('nvlink')), json.dumps(d.get('roofline_nvlink')), json.dumps(d.get('probe_round')), d['roofline_gemm']['achieved'], d['roofline_gemm']['classes']['uniform_philox']['ours_over_cublas'])
PY
tail -4 gpurun_out/bench_n8.err
