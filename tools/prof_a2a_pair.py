"""ncu target for the NVLink wire counters (nvltx__/nvlrx__ bytes, user vs protocol): isolated pairs, one kernel at a time
(ncu serialises kernels, so only one-way traffic can be profiled): push (bulk stores) then pull (bulk loads), 64 MiB."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k3s_nvidia_b200 import _lib as L
from k3s_nvidia_b200.probe import Probe

p = Probe()
for variant in (L.A2A_PUSH_TMA, L.A2A_PULL_TMA):
    r = p.nvlink_a2a([0, 1], bytes_per_pair=64 << 20, mode=L.A2A_PEER_PAIR, warmup=0, reps=1, verify=1, variant=variant)
    print(variant, r.min_pair_gbs, r.max_pair_gbs, r.verified)
p.a2a_release()
