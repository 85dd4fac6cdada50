"""The CPU oracle pinned against the committed golden vectors (tests/golden/golden.json, generated
by tests/golden/make_golden.py from an independent pure-Python restatement).  The reference holds
no vectors for this path ("parity unpinned") — this is the strongest pin available."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import _oracle

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


@pytest.fixture(scope="module")
def o():
    return _oracle.load()


def test_pattern_words(o):
    for i, seed, want in G["pattern"]:
        assert o.oracle_pattern_word(i, seed) == want


def test_pattern_fill_matches_words(o):
    a = _oracle.pattern(o, 2**32 - 3, 8, 0xB200)
    assert [int(v) for v in a] == [o.oracle_pattern_word(2**32 - 3 + j, 0xB200) for j in range(8)]


def test_checksums(o):
    for words, seed, s, x in G["checksum"]:
        assert _oracle.pattern_checksum(o, words, seed) == (s, x)
        if words <= 1 << 20:
            assert _oracle.checksum(o, _oracle.pattern(o, 0, words, seed)) == (s, x)


def test_checksum_is_order_independent_and_additive(o):
    a = _oracle.pattern(o, 0, 4096, 7)
    rng = np.random.default_rng(0)
    b = a.copy()
    rng.shuffle(b)
    assert _oracle.checksum(o, a) == _oracle.checksum(o, b)
    s1, x1 = _oracle.checksum(o, np.ascontiguousarray(a[:1000]))
    s2, x2 = _oracle.checksum(o, np.ascontiguousarray(a[1000:]))
    s, x = _oracle.checksum(o, a)
    assert ((s1 + s2) & (2**64 - 1), x1 ^ x2) == (s, x)


def test_a2a_chunk_seeds(o):
    for seed, src, dst, want in G["a2a_seed"]:
        assert o.oracle_a2a_chunk_seed(seed, src, dst) == want
    seeds = {o.oracle_a2a_chunk_seed(0xB200, s, d) for s in range(8) for d in range(8)}
    assert len(seeds) == 64          # every (src,dst) chunk is distinguishable


def test_gemm_elements_and_dots(o):
    for e, seed, which, k, bits in G["gemm_elem"]:
        assert o.oracle_gemm_elem(e, seed, which) == k / 128.0
        assert o.oracle_gemm_elem_bits(e, seed, which) == bits
    for kdim, row, col, acc, val, bits in G["gemm_dot"]:
        got = o.oracle_gemm_dot(kdim, 0xB200, row, col)
        assert got == val                                  # exact in fp64
        assert got * 16384 == acc
        assert o.oracle_bf16_rne(got) == bits


def test_philox_known_answer_vectors(o):
    """The one externally published pin of this oracle: Random123's Philox4x32-10 known-answer vectors."""
    assert len(G["philox_kat"]) == 3
    for ctr, key, out in G["philox_kat"]:
        assert _oracle.philox(o, ctr, key) == out


def test_gemm_uniform_operands_and_dots(o):
    for e, seed, which, bits in G["gemm_uniform_elem"]:
        assert o.oracle_gemm_uniform_bits(e, seed, which) == bits
        v = o.oracle_gemm_uniform_elem(e, seed, which)
        assert -1.0 <= v <= 1.0
    for kdim, row, col, acc in G["gemm_uniform_dot"]:
        assert o.oracle_gemm_uniform_dot(kdim, 0xB200, row, col) == acc      # same fp64 sum in the same order: bit-identical
    # the class really is U(-1,1): mean ~ 0, variance ~ 1/3 over 64 Ki elements
    v = np.array([o.oracle_gemm_uniform_elem(e, 0xB200, 0) for e in range(1 << 16)])
    assert abs(v.mean()) < 0.01 and abs(v.var() - 1.0 / 3.0) < 0.01
    assert o.oracle_gemm_uniform_tol(8192, 30.0) == 30.0 / 256 + 8192 ** 0.5 / 1024


def test_product_operand_generator_matches_oracle():
    """The library's host-side element function (what its own check multiplies) against the oracle, both classes.
    No GPU needed: b200probe_gemm_operand_bits is plain host code."""
    from k3s_nvidia_b200 import _lib

    lib, o = _lib.load(), _oracle.load()
    for e in list(range(64)) + [8191, 8192, 2**26 - 1, 2**34 + 6]:
        for m in (0, 1):
            assert lib.b200probe_gemm_operand_bits(e, 0xB200, m) == o.oracle_gemm_elem_bits(e, 0xB200, m)
            assert lib.b200probe_gemm_operand_bits(e, 0xB200, m | 2) == o.oracle_gemm_uniform_bits(e, 0xB200, m)


def test_bf16_rne(o):
    for f, bits in G["bf16_rne"]:
        assert o.oracle_bf16_rne(f) == bits


def test_host_sweep_port_results(o):
    """The multi-threaded host sweep (the cpu_baseline 'port') produces the same data results."""
    for mode in (1, 2, 4):
        s, x = C.c_uint64(), C.c_uint32()
        dt = o.oracle_host_sweep(1 << 22, 3, mode, 2, 0xB200, C.byref(s), C.byref(x))
        assert dt > 0
        assert (s.value, x.value) == _oracle.pattern_checksum(o, (1 << 22) // 4, 0xB200)


def test_verify_counts_and_locates_faults(o):
    """oracle_verify (the checker of the device verdict pass) against a pure-Python restatement of
    the pattern rule taken from the golden generator, across the 2^32-word boundary."""
    K = 2654435761
    def pat(i, seed):
        return ((i & 0xFFFFFFFF) * K & 0xFFFFFFFF) ^ seed ^ (i >> 32)
    for first_word in (0, 2**32 - 5):
        a = _oracle.pattern(o, first_word, 64, 0xB200)
        assert [int(v) for v in a] == [pat(first_word + j, 0xB200) for j in range(64)]
        assert _oracle.verify(o, a, 0xB200, first_word) == (0, 2**64 - 1)
        b = a.copy()
        for j in (3, 9, 63):
            b[j] ^= np.uint32(1 << (j % 32))
        assert _oracle.verify(o, b, 0xB200, first_word) == (3, first_word + 3)
        assert _oracle.verify(o, a, 0xB201, first_word) == (64, first_word)
