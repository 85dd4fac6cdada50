"""NFD label hand-off: probe results -> features.d file -> gate label."""
import os

import pytest

from k3s_nvidia_b200 import labels as L
from k3s_nvidia_b200.probe import A2aReport, GemmReport, HbmPoint


def pt(nbytes, mode, gbs, verified=1, resident=False):
    return HbmPoint(nbytes, mode, 0, 1.0, 1.0, gbs, gbs, 0, 0, verified, resident)


def test_hbm_verdict_uses_largest_hbm_resident_copy_point():
    th = L.Thresholds(hbm_min_gbs=5900.0)
    pts = [pt(1 << 20, "copy", 900.0, resident=True), pt(1 << 28, "copy", 6400.0), pt(1 << 30, "copy", 6515.0),
           pt(1 << 30, "read", 6890.0), pt(1 << 30, "write", 7024.0)]
    lab = L.hbm_labels({0: pts}, th)
    assert lab["nvidia.com/b200probe.gpu0.hbm-copy-gbs"] == "6515"
    assert lab["nvidia.com/b200probe.gpu0.hbm-read-gbs"] == "6890"
    assert lab["nvidia.com/b200probe.gpu0.hbm-copy-pct-of-nominal"] == "81"
    assert lab["nvidia.com/b200probe.gpu0.hbm-copy-pct-of-measured"] == "99"
    assert lab["nvidia.com/b200probe.hbm-healthy"] == "true"
    # north_star's gate (>= 90% of 8 TB/s) is an env var away and fails this healthy GPU
    assert L.hbm_labels({0: pts}, L.Thresholds(hbm_min_gbs=7200.0))["nvidia.com/b200probe.hbm-healthy"] == "false"
    # a data mismatch is unhealthy regardless of speed
    bad = pts[:-1] + [pt(1 << 30, "write", 7024.0, verified=0)]
    assert L.hbm_labels({0: bad}, th)["nvidia.com/b200probe.gpu0.hbm-healthy"] == "false"
    # cache-resident sizes alone never produce a healthy verdict
    assert L.hbm_labels({0: [pt(1 << 20, "copy", 9000.0, resident=True)]}, th)["nvidia.com/b200probe.hbm-healthy"] == "false"


def test_nvlink_and_gemm_and_gate(tmp_path):
    th = L.Thresholds(hbm_min_gbs=5900.0, nvlink_min_gbs=690.0, gemm_min_tflops=1000.0)
    g = 2
    rep = A2aReport(g, 1.0, 1.0, [750.0, 740.0], [745.0, 745.0], [[0, 750.0], [740.0, 0]], 740.0, 750.0, 1)
    lab = L.nvlink_labels(rep, th)
    assert lab["nvidia.com/b200probe.gpu0.nvlink-to-gpu1-gbs"] == "750"
    assert lab["nvidia.com/b200probe.nvlink-healthy"] == "true"
    cold = A2aReport(g, 1.0, 1.0, [750.0, 300.0], [745.0, 745.0], [[0, 750.0], [300.0, 0]], 300.0, 750.0, 1)
    assert L.nvlink_labels(cold, th)["nvidia.com/b200probe.gpu1.nvlink-healthy"] == "false"
    gm = GemmReport(8192, 8192, 8192, 0.7, 0.7, 1500.0, 1510.0, 0.0, 0.0, 0.0, 1024, 0, 1, 2, 1)
    lab.update(L.gemm_labels({0: gm}, th))
    lab.update(L.hbm_labels({0: [pt(1 << 30, "copy", 6500.0)]}, th))
    lab.update(L.gate_label(lab))
    assert lab["nvidia.com/b200probe.healthy"] == "true"
    path = L.write_feature_file(lab, str(tmp_path))
    assert os.path.basename(path) == "b200probe"
    back = L.parse_feature_file(open(path).read())
    assert back == lab
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".tmp")]
    lab2 = dict(lab)
    lab2.update(L.nvlink_labels(cold, th))
    lab2.update(L.gate_label(lab2))
    assert lab2["nvidia.com/b200probe.healthy"] == "false"


def test_label_syntax_is_enforced():
    assert L.valid_label("nvidia.com/b200probe.gpu0.hbm-copy-gbs", "6515")
    assert not L.valid_label("nvidia.com/b200probe.x", "6515 GB/s")
    assert not L.valid_label("nvidia.com/" + "x" * 64, "1")
    with pytest.raises(ValueError):
        L.render({"nvidia.com/b200probe.bad": "a b"})


def test_runner_publishes_false_when_probes_cannot_run(tmp_path, monkeypatch):
    """GPU-less box: every active probe raises -> labels say unhealthy; nothing hangs, nothing is
    computed on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import _oracle
    from k3s_nvidia_b200.probe import Probe

    monkeypatch.setenv("MOCK_NVML_DEVICES", "2")
    p = Probe(_oracle.MOCK_NVML)
    try:
        r = L.ActiveProbeRunner(p, features_dir=str(tmp_path), interval_s=3600)
        lab = r.run_once()
        assert lab["nvidia.com/b200probe.hbm-healthy"] == "false"
        assert lab["nvidia.com/b200probe.healthy"] == "false"
        assert os.path.exists(tmp_path / "b200probe")
    finally:
        p.close()


def test_passive_nvlink_status_and_labels(monkeypatch):
    """per-link state + fabric health from NVML, correlated with the active matrix."""
    import ctypes as C

    import _oracle
    from k3s_nvidia_b200.probe import Probe

    monkeypatch.setenv("MOCK_NVML_DEVICES", "2")
    monkeypatch.setenv("MOCK_NVML_LINKS_DOWN", "1:5,1:7")
    p = Probe(_oracle.MOCK_NVML)
    try:
        s0, s1 = p.nvlink_passive(0), p.nvlink_passive(1)
        assert (s0["links_total"], s0["links_active"], s0["active_mask"]) == (18, 18, (1 << 18) - 1)
        assert (s1["links_total"], s1["links_active"]) == (18, 16) and not s1["active_mask"] & ((1 << 5) | (1 << 7))
        assert s0["fabric_state"] == 3 and s0["counters_ok"] == 1
        lab = L.nvlink_passive_labels({0: s0, 1: s1})
        assert lab["nvidia.com/b200probe.gpu0.nvlink-links-ok"] == "true"
        assert lab["nvidia.com/b200probe.gpu1.nvlink-links-ok"] == "false"
        assert lab["nvidia.com/b200probe.gpu1.nvlink-links-active"] == "16"
        assert lab["nvidia.com/b200probe.nvlink-links-ok"] == "false"
        mock = C.CDLL(_oracle.MOCK_NVML)
        mock.mock_nvml_add_nvlink_traffic.argtypes = [C.c_int, C.c_ulonglong]
        mock.mock_nvml_add_nvlink_traffic(0, 8 << 20)
        eff = L.wire_efficiency(s0, p.nvlink_passive(0))
        assert abs(eff - 8 / 9) < 1e-9                       # mock: raw = data * 1.125
        assert L.wire_efficiency({"counters_ok": 0}, s0) is None
        # degraded-bandwidth bit of the fabric health mask fails the device
        bad = dict(s0, fabric_health_mask=1)
        assert L.nvlink_passive_labels({0: bad})["nvidia.com/b200probe.gpu0.nvlink-links-ok"] == "false"
        assert L.nvlink_passive_labels({0: dict(s0, links_total=0, links_active=0)}) == {}
    finally:
        p.close()
