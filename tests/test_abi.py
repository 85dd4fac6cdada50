"""The C-ABI library loads on a GPU-less box, exports every symbol include/b200probe.h declares,
and the NVML half (enumerate + passive health) works against the mock NVML.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

import _oracle

ROOT = _oracle.ROOT
HEADER = os.path.join(ROOT, "include", "b200probe.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200probe_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_and_library_exports_the_same_surface():
    from k3s_nvidia_b200 import _lib

    syms = declared_symbols()
    assert len(syms) >= 29
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200probe.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES drifted from the header"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\b(b200probe_[a-z0-9_]+)\b", out)))
    assert exported == syms, "library exports symbols the header does not declare (or vice versa)"


def test_struct_sizes_match_the_c_compiler(tmp_path):
    """ctypes mirrors vs sizeof() from gcc on the real header."""
    from k3s_nvidia_b200 import _lib

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "b200probe.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(b200probe_device_t),sizeof(b200probe_health_event_t),sizeof(b200probe_hbm_cfg_t),sizeof(b200probe_hbm_result_t),"
                   "sizeof(b200probe_a2a_cfg_t),sizeof(b200probe_a2a_result_t),sizeof(b200probe_gemm_cfg_t),sizeof(b200probe_gemm_result_t));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(t) for t in (_lib.Device, _lib.HealthEvent, _lib.HbmCfg, _lib.HbmResult, _lib.A2aCfg, _lib.A2aResult, _lib.GemmCfg, _lib.GemmResult)]
    assert got == want


def test_strerror_and_uninitialised_calls():
    from k3s_nvidia_b200 import _lib

    lib = _lib.load()
    lib.b200probe_shutdown()
    n = C.c_int()
    assert lib.b200probe_device_count(C.byref(n)) == -2            # ENOTINIT
    for rc in (0, -1, -4, -6, -11, 1002, 2010, 3001, 12345):
        assert lib.b200probe_strerror(rc)
    assert lib.b200probe_pattern_word(5, 0xB200) == ((5 * 2654435761) & 0xFFFFFFFF) ^ 0xB200


def test_init_with_missing_nvml_fails_loudly():
    from k3s_nvidia_b200.probe import Probe, ProbeError

    with pytest.raises(ProbeError) as e:
        Probe("/nonexistent/libnvidia-ml.so.1")
    assert e.value.rc == -3


def test_enumerate_against_mock_nvml(monkeypatch):
    from k3s_nvidia_b200.probe import Probe

    monkeypatch.setenv("MOCK_NVML_DEVICES", "8")
    p = Probe(_oracle.MOCK_NVML)
    try:
        assert p.device_count() == 8
        devs, usec = p.enumerate()
        assert [d.index for d in devs] == list(range(8))
        assert devs[3].uuid == "GPU-b2000000-0000-4000-8000-000000000003"
        assert devs[0].name == "NVIDIA B200" and devs[0].cc == (10, 0) and devs[0].mem_total == 192265846784
        assert devs[0].numa_node == -1 and devs[0].mig_enabled == 0 and devs[0].cuda_ordinal == -1
        assert usec >= 0
    finally:
        p.close()


def test_probe_calls_without_cuda_do_not_fall_back(monkeypatch):
    """On this GPU-less box every probe entry must fail with ENOCUDA — never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from k3s_nvidia_b200.probe import Probe, ProbeError

    monkeypatch.setenv("MOCK_NVML_DEVICES", "1")
    p = Probe(_oracle.MOCK_NVML)
    try:
        with pytest.raises(ProbeError) as e:
            p.hbm_sweep(0, min_bytes=1 << 20, max_bytes=1 << 20)
        assert e.value.rc == -4
        with pytest.raises(ProbeError) as e:
            p.hbm_fill(0, 0x1000, 1024, 1)
        assert e.value.rc == -4
        with pytest.raises(ProbeError):
            p.nvlink_a2a([0, 1], bytes_per_pair=1 << 20)
        with pytest.raises(ProbeError):
            p.gemm(0, m=256, n=256, k=64)
    finally:
        p.close()


def test_missing_library_raises(monkeypatch, tmp_path):
    from k3s_nvidia_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libb200probe.so"))
    with pytest.raises(ImportError):
        _lib.load()
