#!/bin/bash
# compute-sanitizer over the kernels added this session: ring verify mode, staggered/synchronised/mixed exchange, host pipeline
mkdir -p gpurun_out
export B200PROBE_A2A_SYNC_TIMEOUT_US=20000000     # the tools slow kernels down by orders of magnitude
( timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_hbm.py -q -x -k "verdict or pinned or (fill_copy_read and variant0 and 4100)" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|error" | head -20 ) | tee gpurun_out/sanitizer_memcheck_r01b.txt
( timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_a2a.py -q -x -k "exchange_lands and (sync or stagger or mix or auto) and (4112 or 1048576)" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|error" | head -20 ) | tee -a gpurun_out/sanitizer_memcheck_r01b.txt
( timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_hbm.py tests/test_gpu_a2a.py -q -x -k "(verdict and tma) or (exchange_lands and (sync or stagger) and 4112)" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|hazard|error" | head -20 ) | tee gpurun_out/sanitizer_racecheck_r01b.txt
