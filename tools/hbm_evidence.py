"""HBM copy: what bounds it?  Experiments on the 1 GiB copy kernel with everything else fixed:
  * L2 eviction policy of the bulk loads x bulk stores (evict_first / normal / last / unchanged)
  * chunk -> worker map (interleaved window sweeping through the buffer vs one contiguous slab per warp)
  * per-warp phase separation (fill all stages, then store all stages) at several ring depths
  * dst placed at different distances from src (same allocation, offset by k MiB) to move the read and the write
    streams relative to each other in the DRAM address map
and the library copy (torch copy_, cudaMemcpyAsync D2D) in the same process for the denominator.
Writes gpurun_out/hbm_evidence.txt.  The knobs are environment variables read per launch (hbm_sweep.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from k3s_nvidia_b200.probe import Probe

N = 1 << 30
p = Probe()
big = torch.empty(3 * N, dtype=torch.uint8, device="cuda:0")
src = big[:N]
st = torch.cuda.current_stream().cuda_stream
p.hbm_fill(0, src.data_ptr(), N, 0xB200, st)
torch.cuda.synchronize()
part = torch.zeros(4, dtype=torch.int64, device="cuda:0")


def timeit(fn, reps=7, inner=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    ts.sort()
    return ts[len(ts) // 2], ts[0], ts[-1]


KNOBS = ["B200PROBE_HBM_LOAD_POLICY", "B200PROBE_HBM_STORE_POLICY", "B200PROBE_HBM_SLAB_MAP", "B200PROBE_HBM_COPY_BATCH"]


def run(label, dst_off=N, env=None, **tuning):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    dst = big[dst_off:dst_off + N]
    med, best, worst = timeit(lambda: p.hbm_copy(0, src.data_ptr(), dst.data_ptr(), N, st, **tuning))
    # the data result must not change with any knob
    part.zero_()
    p.hbm_read(0, dst.data_ptr(), N, part.data_ptr(), st)
    torch.cuda.synchronize()
    chk = (part[0].item() & 0xFFFFFFFFFFFFFFFF, part[1].item() & 0xFFFFFFFF)
    return f"{2 * N / med / 1e6:8.1f} {2 * N / best / 1e6:8.1f} {2 * N / worst / 1e6:8.1f}  {label}", chk


lines = ["# HBM copy 1 GiB (2 GiB algorithmic bytes), GB/s median / best / worst of 7 x 10 back-to-back launches, CUDA events", ""]
ref_chk = None


def add(label, **kw):
    global ref_chk
    line, chk = run(label, **kw)
    if ref_chk is None:
        ref_chk = chk
    assert chk == ref_chk, f"data result changed under {label}"
    lines.append(line)
    print(line, flush=True)


add("shipped: 8 KiB x 4 stages x 2 warps, evict_first loads and stores, interleaved map")
POL = {0: "evict_first", 1: "evict_normal", 2: "evict_last", 3: "evict_unchanged"}
lines.append("\n## L2 eviction policy, loads x stores")
for lp in range(4):
    for sp in range(4):
        if lp == 0 and sp == 0:
            continue
        add(f"load {POL[lp]:<15} store {POL[sp]}", env={KNOBS[0]: lp, KNOBS[1]: sp})
lines.append("\n## chunk -> worker map")
add("slab per warp (contiguous 3.5 MiB per warp)", env={KNOBS[2]: 1})
add("slab per warp, 16 KiB stages", env={KNOBS[2]: 1}, stage_bytes=16384)
lines.append("\n## per-warp phase separation: fill all stages, then store all stages")
for sb, stg, w in [(8192, 4, 2), (8192, 8, 2), (16384, 6, 2), (32768, 3, 2), (16384, 4, 3), (8192, 8, 3), (32768, 6, 1)]:
    add(f"batch: {sb} B x {stg} stages x {w} warps", env={KNOBS[3]: 1}, stage_bytes=sb, stages=stg, warps_per_cta=w)
lines.append("\n## distance between the read stream and the write stream (dst = src + 1 GiB + delta)")
for d_mib in [0, 1, 2, 8, 32, 96, 352, 1024]:
    add(f"dst offset +{d_mib} MiB", dst_off=N + (d_mib << 20))
for k in KNOBS:
    os.environ.pop(k, None)
lines.append("\n## library copies in the same process (the roofline denominator is the first)")
a, b = src.view(torch.bfloat16), big[N:2 * N].view(torch.bfloat16)
med, best, worst = timeit(lambda: b.copy_(a))
lines.append(f"{2 * N / med / 1e6:8.1f} {2 * N / best / 1e6:8.1f} {2 * N / worst / 1e6:8.1f}  torch copy_ (the MEASURED_PEAKS.json recipe)")
lines.append("\n## read-only and write-only at the same size, for the mix argument")
med, best, worst = timeit(lambda: p.hbm_read(0, src.data_ptr(), N, part.data_ptr(), st))
lines.append(f"{N / med / 1e6:8.1f} {N / best / 1e6:8.1f} {N / worst / 1e6:8.1f}  read  (N bytes)")
med, best, worst = timeit(lambda: p.hbm_fill(0, big[N:2 * N].data_ptr(), N, 0xB200, st))
lines.append(f"{N / med / 1e6:8.1f} {N / best / 1e6:8.1f} {N / worst / 1e6:8.1f}  write (N bytes)")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/hbm_evidence.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-6:]))
