"""Hypothesis check: staggered all-to-all with the steps synchronised (host-side sync between steps).
Per step t every rank r pushes its whole chunk to (r+t) mod G with all its CTAs; each step's kernels
are event-timed per device.  If every step runs near the 2-GPU bidirectional rate (692 GB/s) then a
device-side step barrier is worth building."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe()
g = torch.cuda.device_count()
S = 256 << 20
SEED = 0xB200
ords = (C.c_int * g)(*range(g))
p._check(p.lib.b200probe_enable_peer_access(ords, g), "peer")
wins = [torch.zeros(2 * g * S, dtype=torch.uint8, device=f"cuda:{r}") for r in range(g)]
peers = (C.c_void_p * g)(*[w.data_ptr() for w in wins])
streams = [torch.cuda.current_stream(r).cuda_stream for r in range(g)]
def sync():
    for r in range(g):
        torch.cuda.synchronize(r)
def step(t, variant=2):
    evs = []
    for r in range(g):
        with torch.cuda.device(r):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            p._check(p.lib.b200probe_a2a_exchange(r, r, g, peers, S, SEED, variant, 0, (r + t) % g, streams[r]), "x")
            e1.record()
            evs.append((e0, e1))
    sync()
    return [S / e0.elapsed_time(e1) / 1e6 for e0, e1 in evs]
out = []
for rep in range(3):
    tot = 0.0
    for t in range(1, g):
        rates = step(t)
        tot += S / min(rates) / 1e6
        if rep == 2:
            out.append(f"step t={t} per-device GB/s {[round(x) for x in rates]}")
    out.append(f"rep {rep}: sum of per-step max times {tot:.4f} ms -> {(g-1)*S/tot/1e6:.1f} GB/s per direction (steps synchronised)")
# symmetric pairing (r <-> r^t, XOR schedule): each pair is a closed bidirectional 2-GPU exchange
for rep in range(2):
    tot = 0.0
    for t in range(1, g):
        evs = []
        for r in range(g):
            with torch.cuda.device(r):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                p._check(p.lib.b200probe_a2a_exchange(r, r, g, peers, S, SEED, 2, 0, r ^ t, streams[r]), "x")
                e1.record()
                evs.append((e0, e1))
        sync()
        rates = [S / e0.elapsed_time(e1) / 1e6 for e0, e1 in evs]
        tot += S / min(rates) / 1e6
        if rep == 1:
            out.append(f"xor step t={t} per-device GB/s {[round(x) for x in rates]}")
    out.append(f"xor rep {rep}: {(g-1)*S/tot/1e6:.1f} GB/s per direction (steps synchronised)")
print("\n".join(out))
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/a2a_step_sync_g{g}.txt", "w").write("\n".join(out) + "\n")
