"""Generates tests/golden/golden.json — run here (the container that has /root/reference).

Two kinds of fixtures, both committed so they travel to the GPU box (which has no /root/reference):
 1. reference INPUTS: the literal text of the reference's config/manifests that the drop-in contract
    says must be accepted unchanged (values.yaml, the resource stanzas of the two manifests), read
    from /root/reference at generation time, with their sha256.
 2. golden VECTORS for the oracle: the synthetic-data contract of SURVEY.md §8d restated a second
    time in pure Python integers / numpy (independent of oracle/oracle.c and of the CUDA code).
    The reference itself has no tests or vectors for this path — "parity unpinned".
"""
import hashlib
import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
M32 = 0xFFFFFFFF


def pat(i, seed):
    return (((i & M32) * 2654435761) & M32) ^ seed ^ (i >> 32)


def mix32(x):
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def chunk_seed(seed, src, dst):
    return seed ^ mix32((src * 251 + dst * 7 + 1) & M32)


def gemm_k(e, seed, which):
    h = mix32((((e & M32) * 0x9E3779B1) & M32) ^ mix32((seed + 0x51ED27 * (which + 1)) & M32) ^ (e >> 32))
    return (h & 0xFF) - 128          # value = k / 128


def bf16_bits_of(f):
    u = struct.unpack("<I", struct.pack("<f", f))[0]
    return u >> 16


def bf16_rne(f):
    u = struct.unpack("<I", struct.pack("<f", np.float32(f)))[0]
    r = 0x7FFF + ((u >> 16) & 1)
    return ((u + r) >> 16) & 0xFFFF


# Known-answer vectors of Philox4x32-10 as published with Random123 (kat_vectors: "philox4x32 10 ctr key -> out").
# These are the one EXTERNAL pin of this fixture: the generator below is checked against them before it is used.
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([M32, M32, M32, M32], [M32, M32], [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def philox4x32_10(ctr, key):
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & M32, (p0 >> 32) ^ c[3] ^ k[1], p0 & M32]
        k = [(k[0] + 0x9E3779B9) & M32, (k[1] + 0xBB67AE85) & M32]
    return c


def uniform_bits(e, seed, which):
    """SURVEY.md §8d GEMM probe operands: bf16 U(-1,1) from Philox under key (seed, 0)."""
    r = philox4x32_10([(e // 4) & M32, (e // 4) >> 32, which, 0], [seed, 0])[e % 4]
    x = (r >> 8) / 8388608.0 - 1.0                      # exact
    return bf16_rne(x)


def bf16_value(bits):
    return struct.unpack("<f", struct.pack("<I", bits << 16))[0]


def checksum(words, seed):
    i = np.arange(words, dtype=np.uint64)
    w = ((i & np.uint64(M32)) * np.uint64(2654435761)) & np.uint64(M32)
    w ^= np.uint64(seed)
    w ^= i >> np.uint64(32)
    s = int(w.sum(dtype=np.uint64)) if words else 0      # numpy wraps mod 2^64
    x = int(np.bitwise_xor.reduce(w)) if words else 0
    return s, x


def main():
    g = {"reference_inputs": {}, "pattern": [], "checksum": [], "a2a_seed": [], "gemm_elem": [], "gemm_dot": [], "bf16_rne": [],
         "philox_kat": [], "gemm_uniform_elem": [], "gemm_uniform_dot": []}
    for ctr, key, out in PHILOX_KAT:
        assert philox4x32_10(ctr, key) == out, "the generator's Philox does not reproduce the Random123 known-answer vectors"
        g["philox_kat"].append([ctr, key, out])
    for which in (0, 1):
        for e in list(range(12)) + [8191, 8192, 2**26 - 1, 2**34 + 6]:
            g["gemm_uniform_elem"].append([e, 0xB200, which, uniform_bits(e, 0xB200, which)])
    for kdim, row, col in ((64, 0, 0), (64, 1, 2), (256, 127, 255), (2048, 2047, 1), (8192, 4095, 8191)):
        acc = 0.0
        for k in range(kdim):                             # fp64 accumulate, in k order (what oracle_gemm_uniform_dot does)
            acc += bf16_value(uniform_bits(row * kdim + k, 0xB200, 0)) * bf16_value(uniform_bits(col * kdim + k, 0xB200, 1))
        g["gemm_uniform_dot"].append([kdim, row, col, acc])
    for name in ("values.yaml", "nvidia-smi.yaml", "jellyfin.yaml"):
        with open(os.path.join(REF, name), "rb") as f:
            raw = f.read()
        g["reference_inputs"][name] = {"text": raw.decode(), "sha256": hashlib.sha256(raw).hexdigest()}
    for seed in (0xB200, 1, 0xFFFFFFFF, 0):
        for i in list(range(8)) + [2**28 - 1, 2**32 - 1, 2**32, 2**32 + 5, 2**40 + 123]:
            g["pattern"].append([i, seed, pat(i, seed)])
    for seed in (0xB200, 0xDEADBEEF):
        for words in (0, 1, 3, 4, 1024, 65536 + 7, 1 << 20, 1 << 24):
            s, x = checksum(words, seed)
            g["checksum"].append([words, seed, s, x])
    for s_ in range(8):
        for d_ in range(8):
            g["a2a_seed"].append([0xB200, s_, d_, chunk_seed(0xB200, s_, d_)])
    for which in (0, 1):
        for e in list(range(16)) + [8191, 8192, 2**26 - 1, 2**32 + 9]:
            k = gemm_k(e, 0xB200, which)
            g["gemm_elem"].append([e, 0xB200, which, k, bf16_bits_of(k / 128.0)])
    for kdim, row, col in ((64, 0, 0), (64, 1, 2), (256, 127, 255), (8192, 0, 0), (8192, 4095, 8191), (8192, 8191, 17)):
        acc = 0
        for k in range(kdim):
            acc += gemm_k(row * kdim + k, 0xB200, 0) * gemm_k(col * kdim + k, 0xB200, 1)
        val = acc / 16384.0                               # exact: integer / 2^14
        g["gemm_dot"].append([kdim, row, col, acc, val, bf16_rne(val)])
    for f in (0.0, 1.0, -1.0, 1.00390625, 1.01171875, 3.140625, 3.1415927, -2.7182817, 65504.0, 1e-8, 123456.789, -0.0078125):
        g["bf16_rne"].append([float(np.float32(f)), bf16_rne(f)])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden.json:", {k: len(v) for k, v in g.items()})


if __name__ == "__main__":
    main()
