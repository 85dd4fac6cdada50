#!/bin/bash
# round 2, 2-GPU validation: every GPU test (incl. the 58 NVLink ones), smoke, wire counters (NVML + ncu), stability, bench N=2
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_2gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu_2gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_2gpu.log 2>&1; tail -2 gpurun_out/smoke_2gpu.log
timeout 600 python tools/nvlink_counters.py > gpurun_out/nvlink_counters.log 2>&1; tail -12 gpurun_out/nvlink_counters.log
M=gpu__time_duration.sum,nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvltx__bytes_packet_request.sum,nvltx__bytes_packet_response.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,nvlrx__bytes_data_protocol.sum,nvlrx__bytes_packet_request.sum,nvlrx__bytes_packet_response.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'a2a_ring_kernel' -c 4 --csv --log-file gpurun_out/ncu_nvlink_wire_pairs.csv python tools/prof_a2a_pair.py > gpurun_out/prof_a2a_pair.log 2>&1; tail -3 gpurun_out/prof_a2a_pair.log
timeout 900 python tools/stability.py > gpurun_out/stability.log 2>&1; tail -14 gpurun_out/stability.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 3000 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
