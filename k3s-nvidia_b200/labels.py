"""Probe results -> NFD node labels that gate scheduling (SURVEY.md §5 "NFD label hand-off", §8f.1).

GFD (enabled by /root/reference/values.yaml:1-2) hands labels to Node Feature Discovery by writing
``key=value`` lines into a file under /etc/kubernetes/node-feature-discovery/features.d/; NFD's
*local* source turns them into node labels [RECALLED].  The reference hints at label gating in the
commented selector of /root/reference/nvidia-smi.yaml:6-7 (``nvidia.com/gpu.present: "true"``) and
says the plugin "needs these labels for scheduling" (/root/reference/README.md:99).  We write a
SECOND file in the same directory with ``nvidia.com/b200probe.*`` keys, so no chart value changes.

Thresholds.  north_star asks "healthy at >= 90% of 8 TB/s" (7200 GB/s) and ">= 90% of 900 GB/s/dir".
The driver-measured library copy peak on this pool is 6565.8 GB/s = 82% of 8 TB/s
(MEASURED_PEAKS.json; the part's bus is 7680 bit x 3996 MHz x 2 = 7672 GB/s), so a 7200 GB/s gate
would fail every healthy B200.  Defaults are therefore 90% of the pool-measured healthy figures;
both fractions (of nominal and of measured) are always published, and the north_star gate is one
env var away (B200PROBE_HBM_MIN_GBS=7200).
"""
from __future__ import annotations

import logging
import os
import re
import tempfile
import threading
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

log = logging.getLogger("b200probe.labels")

PREFIX = "nvidia.com/b200probe."
FEATURES_DIR = "/etc/kubernetes/node-feature-discovery/features.d"
FEATURE_FILE = "b200probe"

HBM_NOMINAL_GBS = 8000.0         # north_star's denominator
HBM_MEASURED_GBS = 6565.8        # MEASURED_PEAKS.json hbm_gbs (torch copy_, read+write bytes)
NVLINK_NOMINAL_GBS = 900.0       # per direction per GPU
NVLINK_MEASURED_GBS = 770.0      # /opt/skills/guides/B200_PROFILING.md: peer copy per direction
NVLINK_HEALTHY_GBS = 692.0       # this repo's push kernel with all 18 links up, G = 2 (profiles/a2a_tune_r01_2gpu.txt, +-0.3%)
NVLINK_HEALTHY_GBS_BOX = 700.0   # G > 2: one peer per step with the step barrier (profiles/a2a_sync_relaxed_r01_g8.txt: 701 at every GPU)
GEMM_NOMINAL_TFLOPS = 2250.0
GEMM_MEASURED_TFLOPS = 1670.2

_LABEL_VALUE = re.compile(r"^(([A-Za-z0-9][-A-Za-z0-9_.]*)?[A-Za-z0-9])?$")
_LABEL_NAME = re.compile(r"^([A-Za-z0-9][-A-Za-z0-9_.]*)?[A-Za-z0-9]$")


def _env_float(name: str, default: float) -> float:
    try:
        return float(os.environ[name])
    except (KeyError, ValueError):
        return default


@dataclass
class Thresholds:
    hbm_min_gbs: float = field(default_factory=lambda: _env_float("B200PROBE_HBM_MIN_GBS", 0.90 * HBM_MEASURED_GBS))
    # one dead link of 18 costs 5.6% (692 -> ~654 GB/s): the gate sits 3% under the healthy figure for the
    # number of GPUs exchanged (0 = derive from G when the labels are made; B200PROBE_NVLINK_MIN_GBS pins it)
    nvlink_min_gbs: float = field(default_factory=lambda: _env_float("B200PROBE_NVLINK_MIN_GBS", 0.0))
    gemm_min_tflops: float = field(default_factory=lambda: _env_float("B200PROBE_GEMM_MIN_TFLOPS", 0.70 * GEMM_MEASURED_TFLOPS))
    verdict_min_bytes: int = 256 << 20        # sizes below L2 are cache-resident: never used for the verdict


def valid_label(key: str, value: str) -> bool:
    if "/" in key:
        pfx, name = key.split("/", 1)
        if len(pfx) > 253:
            return False
    else:
        name = key
    return len(name) <= 63 and bool(_LABEL_NAME.match(name)) and len(value) <= 63 and bool(_LABEL_VALUE.match(value))


def _b(x: bool) -> str:
    return "true" if x else "false"


def hbm_labels(per_gpu: Dict[int, list], th: Thresholds) -> Dict[str, str]:
    """per_gpu: NVML index -> list of HbmPoint.  Verdict = copy GB/s at the largest HBM-resident size."""
    out: Dict[str, str] = {}
    all_ok = True
    worst = None
    for idx, pts in sorted(per_gpu.items()):
        big = [p for p in pts if not p.cache_resident and p.bytes >= th.verdict_min_bytes]
        best = {}
        for mode in ("read", "write", "copy"):
            cand = [p for p in big if p.mode == mode]
            if cand:
                best[mode] = max(cand, key=lambda p: p.bytes)
        data_ok = all(p.verified != 0 for p in pts)
        for mode, p in best.items():
            out[f"{PREFIX}gpu{idx}.hbm-{mode}-gbs"] = str(int(round(p.gbs_median)))
        ok = data_ok and "copy" in best and best["copy"].gbs_median >= th.hbm_min_gbs
        if "copy" in best:
            g = best["copy"].gbs_median
            out[f"{PREFIX}gpu{idx}.hbm-copy-pct-of-nominal"] = str(int(round(100.0 * g / HBM_NOMINAL_GBS)))
            out[f"{PREFIX}gpu{idx}.hbm-copy-pct-of-measured"] = str(int(round(100.0 * g / HBM_MEASURED_GBS)))
            worst = g if worst is None else min(worst, g)
        out[f"{PREFIX}gpu{idx}.hbm-data-ok"] = _b(data_ok)
        out[f"{PREFIX}gpu{idx}.hbm-healthy"] = _b(ok)
        all_ok = all_ok and ok
    out[f"{PREFIX}hbm-healthy"] = _b(all_ok and bool(per_gpu))
    if worst is not None:
        out[f"{PREFIX}hbm-copy-min-gbs"] = str(int(round(worst)))
    return out


def nvlink_labels(rep, th: Thresholds) -> Dict[str, str]:
    """rep: probe.A2aReport.  "Per-link" through NVSwitch means per (src,dst) pair:
    published per GPU as egress/ingress GB/s plus the cold-spot of the pair matrix."""
    out: Dict[str, str] = {}
    ok = rep.verified != 0
    min_gbs = th.nvlink_min_gbs or (0.97 * NVLINK_HEALTHY_GBS if rep.g <= 2 else 0.96 * NVLINK_HEALTHY_GBS_BOX)
    for g in range(rep.g):
        out[f"{PREFIX}gpu{g}.nvlink-egress-gbs"] = str(int(round(rep.egress_gbs[g])))
        out[f"{PREFIX}gpu{g}.nvlink-ingress-gbs"] = str(int(round(rep.ingress_gbs[g])))
        good = rep.egress_gbs[g] >= min_gbs
        out[f"{PREFIX}gpu{g}.nvlink-healthy"] = _b(good and rep.verified != 0)
        ok = ok and good
        for p in range(rep.g):
            if p != g and rep.pair_gbs[g][p] > 0:
                out[f"{PREFIX}gpu{g}.nvlink-to-gpu{p}-gbs"] = str(int(round(rep.pair_gbs[g][p])))
    out[f"{PREFIX}nvlink-min-pair-gbs"] = str(int(round(rep.min_pair_gbs)))
    out[f"{PREFIX}nvlink-egress-pct-of-nominal"] = str(int(round(100.0 * min(rep.egress_gbs[: rep.g]) / NVLINK_NOMINAL_GBS)))
    out[f"{PREFIX}nvlink-data-ok"] = _b(rep.verified != 0)
    out[f"{PREFIX}nvlink-healthy"] = _b(ok)
    return out


def nvlink_passive_labels(per_gpu: Dict[int, dict], expected_links: int = 18) -> Dict[str, str]:
    """per_gpu: NVML index -> Probe.nvlink_passive() dict.  Every link NVML knows must be up and the
    fabric registration completed with a clean health mask."""
    out: Dict[str, str] = {}
    all_ok = True
    for idx, st in sorted(per_gpu.items()):
        if st["links_total"] == 0:
            continue                                   # no NVLink on this part: nothing to assert
        ok = st["links_active"] == st["links_total"] and st["links_total"] >= expected_links
        # fabric: a completed registration must have succeeded, and the 2-bit DEGRADED_BW field of the
        # health mask (nvml.h:3453) must not read TRUE (1).  NOT_SUPPORTED / not started is not a fault.
        if st["fabric_state"] == 3 and st["fabric_status"] != 0:
            ok = False
        if (st["fabric_health_mask"] & 0x3) == 1:
            ok = False
        out[f"{PREFIX}gpu{idx}.nvlink-links-active"] = str(st["links_active"])
        out[f"{PREFIX}gpu{idx}.nvlink-links-total"] = str(st["links_total"])
        out[f"{PREFIX}gpu{idx}.nvlink-links-ok"] = _b(ok)
        all_ok = all_ok and ok
    if out:
        out[f"{PREFIX}nvlink-links-ok"] = _b(all_ok)
    return out


def wire_efficiency(before: dict, after: dict):
    """payload / raw bytes on the wire between two passive snapshots (None when the counters are not
    readable): NVML's DATA vs RAW counters quantify the protocol overhead of the active exchange."""
    if not (before.get("counters_ok") and after.get("counters_ok")):
        return None
    d = after["data_tx_kib"] - before["data_tx_kib"]
    r = after["raw_tx_kib"] - before["raw_tx_kib"]
    return (d / r) if r > 0 else None


def gemm_labels(per_gpu: Dict[int, object], th: Thresholds) -> Dict[str, str]:
    out: Dict[str, str] = {}
    all_ok = True
    for idx, r in sorted(per_gpu.items()):
        ok = r.verified == 1 and r.tflops_median >= th.gemm_min_tflops
        out[f"{PREFIX}gpu{idx}.gemm-tflops"] = str(int(round(r.tflops_median)))
        out[f"{PREFIX}gpu{idx}.gemm-data-ok"] = _b(r.verified == 1)
        out[f"{PREFIX}gpu{idx}.gemm-healthy"] = _b(ok)
        all_ok = all_ok and ok
    out[f"{PREFIX}gemm-healthy"] = _b(all_ok and bool(per_gpu))
    return out


def gate_label(labels: Dict[str, str]) -> Dict[str, str]:
    """The one label manifests select on: every probe that ran is healthy."""
    parts = [v for k, v in labels.items() if k in (f"{PREFIX}hbm-healthy", f"{PREFIX}nvlink-healthy", f"{PREFIX}gemm-healthy")]
    return {f"{PREFIX}healthy": _b(bool(parts) and all(v == "true" for v in parts))}


def render(labels: Dict[str, str]) -> str:
    bad = [(k, v) for k, v in labels.items() if not valid_label(k, v)]
    if bad:
        raise ValueError(f"invalid label(s): {bad}")
    return "".join(f"{k}={v}\n" for k, v in sorted(labels.items()))


def write_feature_file(labels: Dict[str, str], features_dir: str = FEATURES_DIR, name: str = FEATURE_FILE) -> str:
    """Atomic replace (NFD may read at any time): write a hidden temp file, then rename."""
    os.makedirs(features_dir, exist_ok=True)
    text = render(labels)
    fd, tmp = tempfile.mkstemp(prefix=".", suffix=".tmp", dir=features_dir)   # dot-files are ignored by NFD
    try:
        with os.fdopen(fd, "w") as f:
            f.write(text)
        os.chmod(tmp, 0o644)
        path = os.path.join(features_dir, name)
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    return path


def parse_feature_file(text: str) -> Dict[str, str]:
    out = {}
    for line in text.splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        k, _, v = line.partition("=")
        out[k] = v
    return out


class ActiveProbeRunner:
    """Runs the active probes on every enumerated GPU at an interval and publishes the labels.
    A probe that raises (timeout, CUDA error, data mismatch) publishes ``…healthy=false``; it never
    blocks ListAndWatch (separate thread, separate CUDA streams)."""

    def __init__(self, probe, *, features_dir: str = FEATURES_DIR, interval_s: float = 600.0, thresholds: Optional[Thresholds] = None,
                 hbm_kwargs: Optional[dict] = None, run_nvlink: bool = True, run_gemm: bool = True):
        self.probe = probe
        self.features_dir = features_dir
        self.interval_s = interval_s
        self.th = thresholds or Thresholds()
        self.hbm_kwargs = hbm_kwargs or dict(min_bytes=1 << 28, max_bytes=1 << 30, warmup=2, reps=5, verify=1)
        self.run_nvlink = run_nvlink
        self.run_gemm = run_gemm
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.last_labels: Dict[str, str] = {}

    def run_once(self) -> Dict[str, str]:
        labels: Dict[str, str] = {}
        n = self.probe.device_count()
        infos = [self.probe.device_info(i) for i in range(n)]
        hbm, gemm = {}, {}
        for d in infos:
            try:
                hbm[d.index] = self.probe.hbm_sweep(d.index, **self.hbm_kwargs)
            except Exception as e:  # noqa: BLE001
                log.error("HBM probe failed on GPU %d: %s", d.index, e)
                labels[f"{PREFIX}gpu{d.index}.hbm-healthy"] = "false"
                labels[f"{PREFIX}hbm-healthy"] = "false"
        if hbm:
            got = hbm_labels(hbm, self.th)
            if labels.get(f"{PREFIX}hbm-healthy") == "false":
                got[f"{PREFIX}hbm-healthy"] = "false"
            labels.update({**got, **{k: v for k, v in labels.items() if v == "false"}})
        if self.run_gemm:
            for d in infos:
                try:
                    gemm[d.index] = self.probe.gemm(d.index, warmup=2, reps=5)
                except Exception as e:  # noqa: BLE001
                    log.error("GEMM probe failed on GPU %d: %s", d.index, e)
                    labels[f"{PREFIX}gpu{d.index}.gemm-healthy"] = "false"
                    labels[f"{PREFIX}gemm-healthy"] = "false"
            if gemm:
                got = gemm_labels(gemm, self.th)
                if labels.get(f"{PREFIX}gemm-healthy") == "false":
                    got[f"{PREFIX}gemm-healthy"] = "false"
                labels.update(got)
        ords = [d.cuda_ordinal for d in [self.probe.device_info(i) for i in range(n)] if d.cuda_ordinal >= 0]
        passive_before = {}
        try:
            passive_before = {d.index: self.probe.nvlink_passive(d.index) for d in infos}
            labels.update(nvlink_passive_labels(passive_before))
        except Exception as e:  # noqa: BLE001
            log.error("passive NVLink status failed: %s", e)
        if self.run_nvlink and len(ords) >= 2:
            try:
                labels.update(nvlink_labels(self.probe.nvlink_a2a(ords, warmup=1, reps=3), self.th))
                if labels.get(f"{PREFIX}nvlink-links-ok") == "false":
                    labels[f"{PREFIX}nvlink-healthy"] = "false"      # a dead link fails the gate even if the matrix still clears the bar
                eff = [wire_efficiency(passive_before[d.index], self.probe.nvlink_passive(d.index)) for d in infos if d.index in passive_before]
                eff = [e for e in eff if e]
                if eff:
                    labels[f"{PREFIX}nvlink-data-over-raw-pct"] = str(int(round(100.0 * min(eff))))
            except Exception as e:  # noqa: BLE001
                log.error("NVLink probe failed: %s", e)
                labels[f"{PREFIX}nvlink-healthy"] = "false"
        labels.update(gate_label(labels))
        labels[f"{PREFIX}timestamp"] = str(int(time.time()))
        self.last_labels = labels
        write_feature_file(labels, self.features_dir)
        return labels

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.run_once()
            except Exception as e:  # noqa: BLE001
                log.error("active probe round failed: %s", e)
            self._stop.wait(self.interval_s)

    def start(self):
        self._thread = threading.Thread(target=self._loop, name="b200probe-active", daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=5.0)
