#!/bin/bash
# round 2: compute-sanitizer over the device code added this round — the 512 x 256 GEMM kernel (both epilogues) and the 256 x 256 /
# single-CTA kernels' tensor-store epilogue, the drained + stamped step barrier of the exchange (needs > 2 GPUs), the copy-engine leg
mkdir -p gpurun_out
export B200PROBE_A2A_SYNC_TIMEOUT_US=20000000     # the tools slow kernels down by orders of magnitude
F='COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|rror'
( timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_gemm.py -q -x -m gpu -k "whole_matrix and (128-256-64 or 256-512-128 or 512-256-192 or 384-256-320)" 2>&1 | grep -E "$F" | head -20 ) | tee gpurun_out/sanitizer_memcheck_r02.txt
# NOTE: the exchange under memcheck must be the KILOBYTE-sized cases (exchange_lands ... 4112): the plugin-entry tests at 8 MiB x 7 exchanges
# x 4 GPUs did not finish in 20 minutes and burnt 86 GPU-minutes (4 GPUs x wall time) in round 2.  Keep every timeout x GPU count in budget.
( timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_a2a.py -q -x -m gpu -k "exchange_lands and (push-sync or auto) and (16 or 4112)" 2>&1 | grep -E "$F" | head -20 ) | tee -a gpurun_out/sanitizer_memcheck_r02.txt
( timeout 240 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_a2a.py -q -x -m gpu -k "(whole_matrix and 256-512-128) or (exchange_lands and push-sync and 4112)" 2>&1 | grep -E "$F" | head -20 ) | tee gpurun_out/sanitizer_racecheck_r02.txt
