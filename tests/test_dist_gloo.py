"""world_size-2 gloo coverage of the N>1 host logic (handle exchange, max-over-ranks timing,
weak-scaling aggregation, unit sharding).  CPU only; the data path itself is per-device."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    from k3s_nvidia_b200 import dist as D
    rank, local_rank, world = D.init("gloo")
    assert world == 2
    # 64-byte IPC-handle-sized payloads are exchanged verbatim, in rank order
    mine = bytes([rank * 16 + (i %% 16) for i in range(64)])
    got = D.all_gather_bytes(mine)
    assert got == [bytes([r * 16 + (i %% 16) for i in range(64)]) for r in range(2)], got
    # timing = max over ranks; value = all units / that time
    agg = D.aggregate_bandwidth(bytes_per_rank=2e9, ms_per_rank=1.0 + rank)     # rank 1 is slower
    assert agg["world"] == 2 and abs(agg["ms"] - 2.0) < 1e-9
    assert abs(agg["gbs"] - 4e9 / 2e-3 / 1e9) < 1e-6
    assert D.reduce_scalar(float(rank), "min") == 0.0 and D.reduce_scalar(float(rank), "sum") == 1.0
    rows = D.all_gather_floats([rank + 0.5, 10.0 * rank])
    assert rows == [[0.5, 0.0], [1.5, 10.0]]
    # independent units shard without overlap and cover everything
    shards = [list(D.shard_units(11, r, 2)) for r in range(2)]
    assert sorted(shards[0] + shards[1]) == list(range(11)) and len(shards[0]) == 6
    D.barrier()
    D.host_barrier()          # on gloo: the plain barrier; on nccl: a CPU-side group, so waiting ranks leave their GPUs idle
    print("rank", rank, "ok")
""") % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(_free_port())
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out}"
        assert f"rank {r} ok" in out


def test_single_process_helpers_without_init():
    from k3s_nvidia_b200 import dist as D

    assert D.all_gather_bytes(b"abc") == [b"abc"]
    assert D.reduce_scalar(3.0) == 3.0
    assert D.aggregate_bandwidth(1e9, 1.0)["gbs"] == 1e9 / 1e-3 / 1e9
    assert list(D.shard_units(5, 0, 1)) == [0, 1, 2, 3, 4]


def test_reference_arm_non_zero_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_prints_exactly_one_json_line_with_the_contract_keys():
    import json

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["metric"].startswith("health-probe HBM GB/s") and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["steps"] >= 1 and d["warmup"] == 3


def test_our_arm_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CPU fallback" in r.stderr
