"""The G x G pair matrix the gpu<i>.nvlink-to-gpu<j>-gbs labels carry (BASELINE config 3), both ways of measuring it:
PEER_ALL (PUSH_SYNC steps, each drained and stamped on the device: one pair per rank at a time, every rank busy) and
PEER_PAIR (one pair alone on an idle fabric, pulls).  Writes gpurun_out/a2a_pair_matrix_g<G>.txt."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from k3s_nvidia_b200 import _lib as L
from k3s_nvidia_b200.probe import Probe

p = Probe()
g = min(torch.cuda.device_count(), 8)
lines = []
for S in (256 << 20, 64 << 20):
    for mode, name in ((L.A2A_PEER_ALL, "PEER_ALL (AUTO)"), (L.A2A_PEER_PAIR, "PEER_PAIR (pull, isolated)")):
        r = p.nvlink_a2a(list(range(g)), bytes_per_pair=S, mode=mode, warmup=2, reps=5, verify=1)
        lines.append(f"# {name}, S = {S >> 20} MiB per pair, {g} GPUs: pair_source={r.pair_source} (0 share, 1 isolated, 2 drained steps)  min {r.min_pair_gbs:.1f}  max {r.max_pair_gbs:.1f}"
                     f"  max/min {r.max_pair_gbs / r.min_pair_gbs:.4f}  verified={r.verified}" + (f"  egress {min(r.egress_gbs[:g]):.1f}-{max(r.egress_gbs[:g]):.1f} GB/s/dir" if mode == 0 else ""))
        for i in range(g):
            lines.append("   " + " ".join(f"{r.pair_gbs[i][j]:7.1f}" for j in range(g)))
p.a2a_release()
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/a2a_pair_matrix_g{g}.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
