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


# ---- the runner against a scripted probe: tenants, memory pressure, staleness, calibration, localisation -------------------
class FakeProbe:
    """Duck-typed stand-in for probe.Probe: canned results per GPU, scripted busy / out-of-memory states, call log."""

    def __init__(self, n=2, copy_gbs=6600.0, egress=700.0):
        from k3s_nvidia_b200.probe import DeviceInfo

        self.infos = [DeviceInfo(i, f"GPU-fake-{i}", "NVIDIA B200", f"0000:{i}", 192 << 30, (10, 0), -1, 0, 0xFF, i) for i in range(n)]
        self.busy, self.nomem = set(), set()
        self.copy_gbs = {i: copy_gbs for i in range(n)}
        self.egress = egress
        self.pair_override = {}
        self.links_down = {}
        self.calls, self.released = [], 0

    def device_count(self):
        return len(self.infos)

    def device_info(self, i):
        return self.infos[i]

    def device_busy(self, i):
        return {"busy": int(i in self.busy), "compute_procs": int(i in self.busy), "util_gpu_pct": 0, "util_mem_pct": 0, "mem_used": 0}

    def hbm_sweep(self, idx, **kw):
        from k3s_nvidia_b200.probe import ProbeError

        self.calls.append(("hbm", idx))
        if idx in self.nomem:
            raise ProbeError(L.ENOMEM, "hbm_sweep", "out of device memory")
        return [pt(1 << 30, "copy", self.copy_gbs[idx]), pt(1 << 30, "read", 7000.0), pt(1 << 30, "write", 6900.0)]

    def gemm(self, idx, **kw):
        self.calls.append(("gemm", idx))
        return GemmReport(8192, 8192, 8192, 0.7, 0.7, 1600.0, 1610.0, 0.0, 0.0, 0.0, 1024, 0, 1, 2, 1)

    def nvlink_passive(self, idx):
        down = self.links_down.get(idx, 0)
        return {"links_total": 18, "links_active": 18 - bin(down).count("1"), "active_mask": ((1 << 18) - 1) & ~down, "fabric_state": 3, "fabric_status": 0,
                "fabric_health_mask": 0, "data_tx_kib": 0, "data_rx_kib": 0, "raw_tx_kib": 0, "raw_rx_kib": 0, "counters_ok": 0}

    def nvlink_a2a(self, ords, **kw):
        from k3s_nvidia_b200 import _lib

        self.calls.append(("a2a", tuple(ords)))
        g = len(ords)
        m = [[0.0 if i == j else self.pair_override.get((ords[i], ords[j]), self.egress) for j in range(g)] for i in range(g)]
        flat = [m[i][j] for i in range(g) for j in range(g) if i != j]
        return A2aReport(g, 1.0, 1.0, [self.egress] * g, [self.egress] * g, m, min(flat), max(flat), 1, _lib.PAIR_STEPPED if g > 2 else _lib.PAIR_ISOLATED)

    def release(self, ords=()):
        self.released += 1


P = L.PREFIX


def test_busy_gpu_is_skipped_and_keeps_its_last_idle_verdict(tmp_path):
    """a GPU carrying tenant work must not be probed (the probe would steal its bandwidth and read
    low against idle-box gates) and must not be published unhealthy; its last idle verdict stands."""
    fp = FakeProbe(n=2)
    r = L.ActiveProbeRunner(fp, features_dir=str(tmp_path), interval_s=600)
    lab = r.run_once()
    assert lab[P + "healthy"] == "true" and lab[P + "gpu1.probe-state"] == "probed" and lab[P + "gpu1.hbm-copy-gbs"] == "6600"
    assert fp.released == 1, "arenas are released after every round"
    fp.busy = {1}
    fp.copy_gbs[1] = 1000.0                      # what a probe under tenant load WOULD read: must never be measured
    fp.calls.clear()
    lab = r.run_once()
    assert ("hbm", 1) not in fp.calls and ("gemm", 1) not in fp.calls and not [c for c in fp.calls if c[0] == "a2a"]
    assert lab[P + "gpu1.probe-state"] == "busy" and lab[P + "gpu0.probe-state"] == "probed"
    assert lab[P + "gpu1.hbm-copy-gbs"] == "6600" and lab[P + "gpu1.hbm-healthy"] == "true"      # carried over
    assert lab[P + "gpu0.nvlink-egress-gbs"] == "700" and lab[P + "nvlink-healthy"] == "true"     # last measured NVLink picture stands
    assert lab[P + "healthy"] == "true"
    # out of device memory is a resource verdict, not a fault
    fp.busy, fp.nomem = set(), {0}
    fp.copy_gbs[1] = 6600.0
    lab = r.run_once()
    assert lab[P + "gpu0.probe-state"] == "no-memory" and lab[P + "gpu0.hbm-healthy"] == "true" and lab[P + "healthy"] == "true"
    # a really slow IDLE gpu still fails
    fp.nomem = set()
    fp.copy_gbs[1] = 1000.0
    lab = r.run_once()
    assert lab[P + "gpu1.hbm-healthy"] == "false" and lab[P + "hbm-healthy"] == "false" and lab[P + "healthy"] == "false"


def test_never_measured_means_no_gate_label(tmp_path):
    fp = FakeProbe(n=2)
    fp.busy = {0, 1}
    r = L.ActiveProbeRunner(fp, features_dir=str(tmp_path), interval_s=600)
    lab = r.run_once()
    assert P + "healthy" not in lab and P + "hbm-healthy" not in lab and not fp.calls
    assert lab[P + "gpu0.probe-state"] == "busy"


def test_feature_file_expires_and_is_withdrawn_on_stop(tmp_path):
    import re
    import time

    fp = FakeProbe(n=1)
    r = L.ActiveProbeRunner(fp, features_dir=str(tmp_path), interval_s=600)
    lab = r.run_once()
    text = open(tmp_path / "b200probe").read()
    m = re.match(r"# \+expiry-time=(\d{4}-\d\d-\d\dT\d\d:\d\d:\d\dZ)\n", text)
    assert m, text[:80]
    import calendar

    exp = calendar.timegm(time.strptime(m.group(1), "%Y-%m-%dT%H:%M:%SZ"))
    assert 1200 <= exp - time.time() <= 1300                   # now + 2 intervals + a minute
    assert L.parse_feature_file(text) == lab and P + "timestamp" not in lab
    r.stop()
    assert not os.path.exists(tmp_path / "b200probe")


def test_gates_follow_the_nodes_own_calibration(tmp_path, monkeypatch):
    monkeypatch.delenv("B200PROBE_HBM_MIN_GBS", raising=False)
    monkeypatch.delenv("B200PROBE_GEMM_MIN_TFLOPS", raising=False)
    fp = FakeProbe(n=1, copy_gbs=7000.0)                       # this node's HBM is faster than the pool figure
    r = L.ActiveProbeRunner(fp, features_dir=str(tmp_path), interval_s=600)
    assert r.run_once()[P + "gpu0.hbm-healthy"] == "true"
    cal = L.parse_feature_file(open(tmp_path / ".b200probe-state" / "calibration").read())
    assert float(cal["hbm-copy-gbs.GPU-fake-0"]) == 7000.0 and float(cal["gemm-tflops.GPU-fake-0"]) == 1600.0
    fp.copy_gbs[0] = 6100.0                                    # clears 0.9 x the pool constant (5909) but not 0.9 x its own 7000
    r2 = L.ActiveProbeRunner(fp, features_dir=str(tmp_path), interval_s=600)      # a restarted daemon reads the file back
    assert r2.run_once()[P + "gpu0.hbm-healthy"] == "false"
    # an implausible first reading is never adopted as the node's own figure
    fp2 = FakeProbe(n=1, copy_gbs=3000.0)
    d2 = tmp_path / "other"
    r3 = L.ActiveProbeRunner(fp2, features_dir=str(d2), interval_s=600)
    assert r3.run_once()[P + "gpu0.hbm-healthy"] == "false"
    assert "hbm-copy-gbs.GPU-fake-0" not in L.parse_feature_file(open(d2 / ".b200probe-state" / "calibration").read())
    # an explicit gate (north_star's literal 7200) overrides calibration
    monkeypatch.setenv("B200PROBE_HBM_MIN_GBS", "7200")
    fp.copy_gbs[0] = 7100.0
    assert L.ActiveProbeRunner(fp, features_dir=str(tmp_path), interval_s=600).run_once()[P + "gpu0.hbm-healthy"] == "false"


def test_cold_cell_is_joined_with_passive_link_state(tmp_path):
    """the active matrix names the cold cell, the passive link state names the suspect and the evidence."""
    fp = FakeProbe(n=4)
    r = L.ActiveProbeRunner(fp, features_dir=str(tmp_path), interval_s=600)
    lab = r.run_once()
    assert lab[P + "nvlink-cold-cell"] == "none" and lab[P + "nvlink-suspect"] == "none" and lab[P + "nvlink-pair-ref-gbs"] == "700"
    # GPU 2 lost two links: its whole row and column read 11 % low, NVML shows the links down
    for q in (0, 1, 3):
        fp.pair_override[(2, q)] = 620.0
        fp.pair_override[(q, 2)] = 621.0
    fp.links_down = {2: (1 << 5) | (1 << 7)}
    lab = r.run_once()
    assert lab[P + "nvlink-cold-cell"] == "gpu2-to-gpu0" and lab[P + "nvlink-cold-cells"] == "6"
    assert lab[P + "nvlink-suspect"] == "gpu2" and lab[P + "nvlink-suspect-evidence"] == "links-down"
    assert lab[P + "gpu2.nvlink-links-down-mask"] == "0xa0"
    assert lab[P + "nvlink-healthy"] == "false"                 # links-ok false fails the gate
    # same cold row without any link reported down: the egress side of GPU 1 is the suspect
    fp.pair_override = {(1, q): 600.0 for q in (0, 2, 3)}
    fp.links_down = {}
    lab = r.run_once()
    assert lab[P + "nvlink-suspect"] == "gpu1" and lab[P + "nvlink-suspect-evidence"] == "egress-cold"
    # a single cold pair stays a pair
    fp.pair_override = {(0, 3): 500.0}
    lab = r.run_once()
    assert lab[P + "nvlink-cold-cell"] == "gpu0-to-gpu3" and lab[P + "nvlink-suspect-evidence"] == "pair-only"
    assert all(L.valid_label(k, v) for k, v in lab.items())


def test_busy_query_through_the_abi_on_mock_nvml(monkeypatch):
    import _oracle
    from k3s_nvidia_b200.probe import Probe

    monkeypatch.setenv("MOCK_NVML_DEVICES", "3")
    monkeypatch.setenv("MOCK_NVML_BUSY", "1:2:0,2:0:35")
    p = Probe(_oracle.MOCK_NVML)
    try:
        b0, b1, b2 = p.device_busy(0), p.device_busy(1), p.device_busy(2)
        assert (b0["busy"], b0["compute_procs"], b0["util_gpu_pct"]) == (0, 0, 0)
        assert (b1["busy"], b1["compute_procs"]) == (1, 2)              # two foreign compute processes
        assert (b2["busy"], b2["compute_procs"], b2["util_gpu_pct"]) == (1, 0, 35)      # nobody listed, but the GPU is working
        monkeypatch.setenv("MOCK_NVML_BUSY", "1:0:9")
        assert p.device_busy(1)["busy"] == 0                           # below B200PROBE_BUSY_UTIL_PCT
    finally:
        p.close()
