"""BASELINE config 4: `values.yaml` time-slicing `replicas: 4` (/root/reference/values.yaml:16-18)
lets four pods share one physical GPU; four concurrent probe processes on ONE B200 must still see
the HBM's bandwidth in aggregate (time-slicing gives no memory or fault isolation).

Each worker process = one "probe pod": it binds the same GPU, keeps its own 2 x S buffers resident,
and launches the copy probe back to back for a common wall-clock window that starts at a shared
barrier.  aggregate GB/s = sum over workers of the algorithmic bytes they moved / the common window.
"""
from __future__ import annotations

import multiprocessing as mp
import time
from typing import Dict, List


def _worker(idx: int, ordinal: int, nbytes: int, window_s: float, barrier, out_q) -> None:
    import torch

    from .probe import Probe

    torch.cuda.set_device(ordinal)
    p = Probe()
    src = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{ordinal}")
    dst = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{ordinal}")
    st = torch.cuda.current_stream().cuda_stream
    p.hbm_fill(ordinal, src.data_ptr(), nbytes, 0xB200 + idx, st)
    for _ in range(3):
        p.hbm_copy(ordinal, src.data_ptr(), dst.data_ptr(), nbytes, st)
    torch.cuda.synchronize()
    barrier.wait()
    t0 = time.perf_counter()
    launches = 0
    batch = 16
    while time.perf_counter() - t0 < window_s:
        for _ in range(batch):
            p.hbm_copy(ordinal, src.data_ptr(), dst.data_ptr(), nbytes, st)
        torch.cuda.synchronize()
        launches += batch
    elapsed = time.perf_counter() - t0
    part = torch.zeros(4, dtype=torch.int64, device=f"cuda:{ordinal}")
    p.hbm_read(ordinal, dst.data_ptr(), nbytes, part.data_ptr(), st)
    torch.cuda.synchronize()
    out_q.put({"worker": idx, "launches": launches, "elapsed_s": elapsed, "bytes": 2.0 * nbytes * launches,
               "sum64": part[0].item() & 0xFFFFFFFFFFFFFFFF, "xor32": part[1].item() & 0xFFFFFFFF})


def run(replicas: int = 4, ordinal: int = 0, nbytes: int = 1 << 30, window_s: float = 5.0) -> Dict:
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(replicas)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(i, ordinal, nbytes, window_s, barrier, q)) for i in range(replicas)]
    for pr in procs:
        pr.start()
    res: List[Dict] = [q.get(timeout=window_s + 180) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    res.sort(key=lambda r: r["worker"])
    window = max(r["elapsed_s"] for r in res)
    total = sum(r["bytes"] for r in res)
    per = [r["bytes"] / r["elapsed_s"] / 1e9 for r in res]
    return {"replicas": replicas, "bytes_per_buffer": nbytes, "window_s": round(window, 3), "aggregate_gbs": round(total / window / 1e9, 1),
            "per_process_gbs": [round(x, 1) for x in per], "spread": round((max(per) - min(per)) / max(per), 4), "workers": res}


if __name__ == "__main__":
    import json
    import sys

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    for size in (256 << 20, 1 << 30):
        print(json.dumps(run(n, 0, size, 5.0)))
