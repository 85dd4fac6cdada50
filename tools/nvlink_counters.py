"""NVLink wire efficiency, counter-backed: NVML's per-device DATA (payload) and RAW (payload + protocol)
throughput counters around each kind of exchange, in ONE process over all visible GPUs, next to the achieved rate:
  two-way  : push (our bulk stores), pull (our bulk loads), mixed, PUSH_SYNC (G > 2), NCCL send/recv, copy engines
  one-way  : isolated pairs, push and pull
payload/raw per direction says how much of the 900 GB/s port a schedule can carry as payload at all; achieved/payload-share
says how close the kernel is to THAT.  Writes gpurun_out/nvlink_counters_g<G>.txt."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from k3s_nvidia_b200 import _lib as L
from k3s_nvidia_b200.probe import Probe

p = Probe()
g = min(torch.cuda.device_count(), 8)
S = int(os.environ.get("A2A_S", str(256 << 20)))
REPS = 20
ords = list(range(g))
p.hbm_sweep(0, min_bytes=1 << 20, max_bytes=1 << 20, warmup=0, reps=1)       # first probe call: CUDA initialised, NVML index <-> CUDA ordinal resolved by UUID
idx_of = {p.device_info(i).cuda_ordinal: i for i in range(p.device_count())}
lines = [f"# NVLink counters, {g} GPUs, S = {S >> 20} MiB per pair, {REPS} timed exchanges + 2 warm-ups per row (PUSH_SYNC: + 3 drained matrix passes)",
         "# GB/s = per direction per GPU, payload bytes / median exchange time; NVML counters are per device, summed over its 18 links, KiB;",
         "# data/raw = payload share of the bytes on the wire in that direction of that device (min and max over the devices)",
         f"{'schedule':<34} {'GB/s':>7} {'tx data/raw':>13} {'rx data/raw':>13} {'tx raw B per payload B':>23} {'data_tx vs algorithmic':>23}"]


def snap():
    return {o: p.nvlink_passive(idx_of[o]) for o in ords}


def run(name, **kw):
    p.a2a_release()
    rep0 = p.nvlink_a2a(ords, bytes_per_pair=S, warmup=1, reps=1, verify=0, **kw)      # context built, first-touch done
    before = snap()
    rep = p.nvlink_a2a(ords, bytes_per_pair=S, warmup=2, reps=REPS, verify=1, **kw)
    after = snap()
    n_ex = REPS + 2 + (3 if rep.pair_source == L.PAIR_STEPPED else 0)
    if kw.get("mode") == L.A2A_PEER_PAIR:
        gbs = f"{rep.min_pair_gbs:.0f}-{rep.max_pair_gbs:.0f}"
        payload_tx = n_ex * (g - 1) * S               # every device sent each of its g-1 chunks n_ex times (one pair at a time)
    else:
        gbs = f"{(g - 1) * S / rep.ms_median / 1e6:.1f}"
        payload_tx = n_ex * (g - 1) * S
    tx, rx, over, alg = [], [], [], []
    for o in ords:
        b, a = before[o], after[o]
        if not (b["counters_ok"] and a["counters_ok"]):
            continue
        dtx, rtx = (a["data_tx_kib"] - b["data_tx_kib"]) * 1024.0, (a["raw_tx_kib"] - b["raw_tx_kib"]) * 1024.0
        drx, rrx = (a["data_rx_kib"] - b["data_rx_kib"]) * 1024.0, (a["raw_rx_kib"] - b["raw_rx_kib"]) * 1024.0
        if rtx > 0:
            tx.append(dtx / rtx)
            over.append(rtx / payload_tx)
            alg.append(dtx / payload_tx)
        if rrx > 0:
            rx.append(drx / rrx)
    f = lambda v: f"{min(v):.3f}-{max(v):.3f}" if v else "n/a"
    line = f"{name:<34} {gbs:>7} {f(tx):>13} {f(rx):>13} {f(over):>23} {f(alg):>23}  verified={rep.verified}"
    lines.append(line)
    print(line, flush=True)


run("two-way push (PUSH_TMA)", mode=L.A2A_PEER_ALL, variant=L.A2A_PUSH_TMA)
run("two-way pull (PULL_TMA)", mode=L.A2A_PEER_ALL, variant=L.A2A_PULL_TMA)
run("two-way mixed 50/50 (MIX_TMA)", mode=L.A2A_PEER_ALL, variant=L.A2A_MIX_TMA)
run("two-way push, direct stores", mode=L.A2A_PEER_ALL, variant=L.A2A_PUSH_DIRECT)
if g > 2:
    run("two-way push, stepped (PUSH_SYNC)", mode=L.A2A_PEER_ALL, variant=L.A2A_PUSH_SYNC)
run("two-way copy engines (memcpyPeer)", mode=L.A2A_CE)
run("two-way NCCL send/recv", mode=L.A2A_NCCL)
run("one-way push, isolated pairs", mode=L.A2A_PEER_PAIR, variant=L.A2A_PUSH_TMA)
run("one-way pull, isolated pairs", mode=L.A2A_PEER_PAIR, variant=L.A2A_PULL_TMA)
p.a2a_release()
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/nvlink_counters_g{g}.txt", "w").write("\n".join(lines) + "\n")
