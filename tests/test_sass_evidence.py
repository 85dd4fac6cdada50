"""What proves a Blackwell-native kernel (B200_PROFILING.md): the SASS of the built library, per kernel.  CPU-only —
`cuobjdump` reads the cubin embedded in libb200probe.so; skipped where the CUDA toolkit is not installed."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "k3s-nvidia_b200", "libb200probe.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, timeout=300).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and "/*" in line:
            kernels[name].append(line)
    assert "arch = sm_100a" in out
    return {k: "\n".join(v) for k, v in kernels.items()}


def _of(sass, needle):
    got = [v for k, v in sass.items() if needle in k]
    assert got, f"no kernel matching {needle}: {sorted(sass)[:20]}"
    return "\n".join(got)


def test_gemm_probe_is_tcgen05_with_tma_and_tmem(sass):
    for kernel in ("gemm_bf16_tn_kernel", "gemm_bf16_tn_2cta_kernel"):
        s = _of(sass, kernel)
        assert "UTCHMMA" in s and "UTMALDG" in s and "LDTM" in s          # tcgen05.mma, TMA tensor loads, tcgen05.ld
        assert "HMMA." not in s.replace("UTCHMMA", "") and "HGMMA" not in s  # no legacy mma.sync / Hopper wgmma path
    assert "UTCHMMA.2CTA" in _of(sass, "gemm_bf16_tn_2cta_kernel")           # cta_group::2
    assert "UTCBAR.2CTA.MULTICAST" in _of(sass, "gemm_bf16_tn_2cta_kernel")  # commit multicast to both CTAs' barriers
    # the 256 x 256 kernels leave through tensor stores; the shipped 512 x 256 kernel has both epilogues compiled in:
    # 32-byte stores straight from registers (default) and the staged tensor store
    assert "UTMASTG" in _of(sass, "gemm_bf16_tn_2cta_kernel") and "UTMASTG" in _of(sass, "gemm_bf16_tn_kernel")
    big = _of(sass, "gemm_bf16_tn_2cta_512_kernel")
    assert "UTCHMMA.2CTA" in big and "UTMALDG" in big and "UTCBAR.2CTA.MULTICAST" in big and "HMMA." not in big.replace("UTCHMMA", "")
    assert "STG.E.ENL2.256" in big and "UTMASTG" in big
    assert big.count("LDTM.x32") >= 16                                       # two register sets of tcgen05.ld in flight, both epilogues


def test_hbm_ring_kernels_move_data_with_bulk_tma(sass):
    s = _of(sass, "hbm_ring_kernel")
    assert "UBLKCP.S.G" in s and "UBLKCP.G.S" in s                        # cp.async.bulk global->shared and shared->global
    assert "SYNCS.ARRIVE.TRANS64" in s and "TRYWAIT" in s                  # mbarrier expect_tx / try_wait
    copy = _of(sass, "hbm_ring_kernelILi4E")                               # COPY: the bulk path never touches registers —
    assert not re.search(r"LD[GS]\.(E\.)?128|ST[GS]\.(E\.)?128|\bLDS\b|\bSTS\b", copy)   # only 4-byte LDG/STG of the <16-byte tail


def test_nvlink_exchange_kernels_store_with_bulk_tma_and_signal_relaxed(sass):
    ring, stagger = _of(sass, "a2a_ring_kernel"), _of(sass, "a2a_stagger_kernel")
    assert "UBLKCP.G.S" in ring and "UBLKCP.S.G" in ring                   # push (stores) and pull (loads) over peer mappings
    assert "UBLKCP.G.S" in stagger
    # the step barrier paces with relaxed system-scope accesses; a release fence there (MEMBAR.SYS before the flag store)
    # drained the bulk-store pipeline at every step (profiles/a2a_sweep_table_r01_g8_fenced_barrier.txt)
    assert "MEMBAR.SC.SYS" not in stagger and "MEMBAR.ALL.SYS" not in stagger
    assert "STG.E.STRONG.SYS" in stagger and "LDG.E.STRONG.SYS" in stagger and "MEMBAR.ALL.GPU" not in stagger


def test_no_kernel_spills(sass):
    for name, body in sass.items():
        assert "STL" not in body and "LDL" not in body, f"{name} spills to local memory"
