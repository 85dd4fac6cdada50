"""GEMM 8192^3: the CTA-pair kernels (variant 3 = 512 x 256 per pair, 2 = 256 x 256 per pair) under their tuning knobs against
cuBLAS on the SAME buffers, both operand classes.  Protocol per row = bench.py's: burst (median of 10 single launches, idle
between) and sustained (back to back for 2 s); cuBLAS is re-measured at the start, in the middle and at the end.
Writes gpurun_out/gemm_tune.txt."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from k3s_nvidia_b200.probe import Probe

p = Probe()
M = N = K = 8192
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
A = torch.empty(M * K, dtype=torch.int16, device=dev)
B = torch.empty(N * K, dtype=torch.int16, device=dev)
Cm = torch.empty(M * N, dtype=torch.int16, device=dev)
Cl = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
Ab, Bb = A.view(torch.bfloat16).view(M, K), B.view(torch.bfloat16).view(N, K)
flop = 2.0 * M * N * K
KNOBS = ["B200PROBE_GEMM_VARIANT", "B200PROBE_GEMM_GROUP_M", "B200PROBE_GEMM_POL_A", "B200PROBE_GEMM_POL_B", "B200PROBE_GEMM_POL_C", "B200PROBE_GEMM_PREFETCH", "B200PROBE_GEMM_EXPT", "B200PROBE_GEMM_EPI", "B200PROBE_GEMM_SKEW"]


def ours():
    p._check(p.lib.b200probe_gemm_launch(0, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, st), "gemm_launch")


def cublas():
    torch.matmul(Ab, Bb.t(), out=Cl)


def burst(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return flop / (statistics.median(ts) * 1e-3) / 1e12


def sustained(fn, seconds=2.0):
    tot, n = 0.0, 0
    while tot < seconds * 1e3:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100):
            fn()
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
        n += 100
    return flop * n / (tot * 1e-3) / 1e12


lines = ["# GEMM 8192^3 bf16, TFLOP/s: burst = median of 10 single launches; sustained = back to back for 2 s (power cap)"]
for cls, cname in ((1, "U(-1,1) Philox operands"), (0, "k/128 operands")):
    p._check(p.lib.b200probe_gemm_fill(0, A.data_ptr(), M * K, 0xB200, 0 | (cls << 1), st), "fill")
    p._check(p.lib.b200probe_gemm_fill(0, B.data_ptr(), N * K, 0xB200, 1 | (cls << 1), st), "fill")
    torch.cuda.synchronize()
    lines.append(f"## {cname}")
    combos = [dict(v=3, g=4), dict(v=2, g=8), "cublas", dict(v=3, g=2), dict(v=3, g=8), dict(v=3, g=1), dict(v=3, g=16), dict(v=3, g=3), "cublas",
              dict(v=3, g=4, c=1), dict(v=3, g=4, a=2), dict(v=3, g=4, b=1), dict(v=3, g=4, a=2, b=1, c=1), dict(v=2, g=8), dict(v=3, g=4), "cublas"]
    if os.environ.get("GEMM_TUNE_QUICK"):
        combos = [dict(v=3, g=4), "cublas", dict(v=3, g=4, s=1), dict(v=3, g=4, s=2), dict(v=3, g=4, e=1), dict(v=3, g=8), "cublas", dict(v=2, g=8), dict(v=3, g=4, x=1), dict(v=3, g=4), "cublas"]
    if os.environ.get("GEMM_TUNE_QUICK") == "prefetch":
        combos = [dict(v=3, g=4), "cublas", dict(v=3, g=4, p=2), dict(v=3, g=4, p=4), dict(v=3, g=4, p=8), dict(v=3, g=4, p=16), "cublas", dict(v=3, g=4, p=32),
                  dict(v=3, g=8, p=8), dict(v=3, g=4), dict(v=3, g=4, p=8), "cublas"]
    for kw in combos:
        for k in KNOBS:
            os.environ.pop(k, None)
        if kw == "cublas":
            line = f"cuBLAS (torch.matmul, same buffers)                      burst {burst(cublas):7.1f}   sustained {sustained(cublas):7.1f}"
        else:
            os.environ[KNOBS[0]] = str(kw["v"])
            os.environ[KNOBS[1]] = str(kw["g"])
            for key, env in (("a", KNOBS[2]), ("b", KNOBS[3]), ("c", KNOBS[4]), ("p", KNOBS[5]), ("x", KNOBS[6]), ("e", KNOBS[7]), ("s", KNOBS[8])):
                if key in kw:
                    os.environ[env] = str(kw[key])
            bu, su = burst(ours), sustained(ours)
            cublas()
            torch.cuda.synchronize()
            same = bool((Cm.view(torch.bfloat16).view(M, N) == Cl).all().item())
            tile = "512x256" if kw["v"] == 3 else "256x256"
            line = (f"ours {tile} band {kw['g']:>2}  A-pol {kw.get('a', 0)} B-pol {kw.get('b', 0)} C-pol {kw.get('c', 0)} prefetch {kw.get('p', 0):>2} expt {kw.get('x', 0)} epi {'tma' if kw.get('e') else 'stg'} skew {kw.get('s', 0)}  burst {bu:7.1f}   sustained {su:7.1f}"
                    f"   C == cuBLAS: {same}")
        lines.append(line)
        print(line, flush=True)
for k in KNOBS:
    os.environ.pop(k, None)
os.makedirs("gpurun_out", exist_ok=True)
open(os.environ.get("GEMM_TUNE_OUT", "gpurun_out/gemm_tune.txt"), "w").write("\n".join(lines) + "\n")
