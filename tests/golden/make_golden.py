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


def checksum(words, seed):
    i = np.arange(words, dtype=np.uint64)
    w = ((i & np.uint64(M32)) * np.uint64(2654435761)) & np.uint64(M32)
    w ^= np.uint64(seed)
    w ^= i >> np.uint64(32)
    s = int(w.sum(dtype=np.uint64)) if words else 0      # numpy wraps mod 2^64
    x = int(np.bitwise_xor.reduce(w)) if words else 0
    return s, x


def main():
    g = {"reference_inputs": {}, "pattern": [], "checksum": [], "a2a_seed": [], "gemm_elem": [], "gemm_dot": [], "bf16_rne": []}
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
