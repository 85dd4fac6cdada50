"""GPU parity for the NVLink all-to-all (needs >= 2 GPUs: `gpurun --gpus 2`).  Bit-exact: every
chunk that lands in a peer window equals the oracle's pattern under the oracle's chunk seed."""
import ctypes as C

import numpy as np
import pytest

import _oracle

pytestmark = pytest.mark.gpu

SEED = 0xB200


@pytest.fixture(scope="module")
def env():
    import torch

    from k3s_nvidia_b200.probe import Probe

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    return torch, Probe(), _oracle.load()


@pytest.mark.parametrize("S", [16, 4096 + 16, 1 << 20, (3 << 20) + 48])
@pytest.mark.parametrize("from_buf", [False, True])
@pytest.mark.parametrize("variant", [0, 1], ids=["tma", "direct"])
def test_push_lands_oracle_pattern_in_every_peer_window(env, S, from_buf, variant):
    torch, p, o = env
    g = min(torch.cuda.device_count(), 8)
    ords = (C.c_int * g)(*range(g))
    p._check(p.lib.b200probe_enable_peer_access(ords, g), "enable_peer_access")
    wins = [torch.full((g * S + 64,), 0x5A, dtype=torch.uint8, device=f"cuda:{r}") for r in range(g)]
    peers = (C.c_void_p * g)(*[w.data_ptr() for w in wins])
    for r in range(g):
        st = torch.cuda.current_stream(r).cuda_stream
        if from_buf:
            send = np.concatenate([_oracle.pattern(o, 0, S // 4, o.oracle_a2a_chunk_seed(SEED, r, d)) for d in range(g)])
            sb = torch.from_numpy(send.view(np.uint8)).to(f"cuda:{r}")
            p._check(p.lib.b200probe_a2a_push_buf(r, r, g, sb.data_ptr(), peers, S, 0, variant, st), "push_buf")
        else:
            p._check(p.lib.b200probe_a2a_push(r, r, g, peers, S, SEED, 3, variant, st), "push")
    for r in range(g):
        torch.cuda.synchronize(r)
    for dst in range(g):
        host = wins[dst].cpu().numpy()
        assert (host[g * S:] == 0x5A).all(), "wrote past the window"
        for src in range(g):
            want = _oracle.pattern(o, 0, S // 4, o.oracle_a2a_chunk_seed(SEED, src, dst))
            got = host[src * S:(src + 1) * S].view(np.uint32)
            assert np.array_equal(got, want), f"chunk {src}->{dst} differs from oracle"
            assert p.lib.b200probe_a2a_chunk_seed(SEED, src, dst) == o.oracle_a2a_chunk_seed(SEED, src, dst)


@pytest.mark.parametrize("variant", [0, 1], ids=["tma", "direct"])
@pytest.mark.parametrize("mode", [0, 1, 2], ids=["peer-all", "peer-pair", "nccl"])
def test_single_process_probe_verifies_and_reports(env, mode, variant):
    torch, p, o = env
    g = min(torch.cuda.device_count(), 8)
    rep = p.nvlink_a2a(list(range(g)), bytes_per_pair=8 << 20, mode=mode, warmup=1, reps=3, verify=1, variant=variant)
    assert rep.verified == 1 and rep.g == g
    for i in range(g):
        for j in range(g):
            assert (rep.pair_gbs[i][j] > 0) == (i != j)
    if mode != 1:
        assert all(x > 0 for x in rep.egress_gbs) and rep.ms_median > 0


def test_full_size_exchange_properties(env):
    """BASELINE config 3 at S = 256 MiB: library-verified landing of all G*(G-1) chunks, and the
    egress figure is stable run to run within 5% (north_star asks 1% for fp results; link timing
    is noisier, so the bound tested here is looser and the measured spread is reported in bench)."""
    torch, p, o = env
    g = min(torch.cuda.device_count(), 8)
    a = p.nvlink_a2a(list(range(g)), bytes_per_pair=256 << 20, warmup=2, reps=5, verify=1)
    b = p.nvlink_a2a(list(range(g)), bytes_per_pair=256 << 20, warmup=2, reps=5, verify=1)
    assert a.verified == 1 and b.verified == 1
    for x, y in zip(a.egress_gbs, b.egress_gbs):
        assert abs(x - y) / max(x, y) < 0.05
