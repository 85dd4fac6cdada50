// Internal helpers shared by the translation units of libb200probe.so.  Not part of the ABI.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../include/b200probe.h"

namespace b200 {

// thread-local detailed error text (b200probe_last_error)
void set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
const char* get_error();

// NVML index -> CUDA ordinal (by UUID); initialises CUDA lazily.  Returns a b200probe rc.
int cuda_ordinal_of(int nvml_idx, int* ordinal);
// Make sure CUDA is usable and `ordinal` is an sm_100 device; caches per-device properties.
struct DevProps {
    int  sms;
    int  l2_bytes;
    int  smem_optin;
    int  cc_major, cc_minor;
};
int device_props(int ordinal, DevProps* out);

// Device-side check of a buffer against the closed-form pattern (hbm_sweep.cu): checksum + count of
// words that differ.  d_partials = 4 x u64 scratch on the device; synchronises `stream`.
struct VerifyResult { uint64_t sum; uint32_t x; uint64_t bad, first; };
int verify_pattern(int ordinal, const void* buf, uint64_t bytes, uint32_t seed, unsigned long long* d_partials, void* stream, VerifyResult* out);

inline int cuda_rc(int cuda_err) { return cuda_err == 0 ? 0 : B200PROBE_CUDA_BASE + cuda_err; }

}  // namespace b200

#define B200_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            b200::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return b200::cuda_rc((int)e__);                                                   \
        }                                                                                     \
    } while (0)

// A failed cudaMalloc is a RESOURCE verdict (a tenant holds the memory), not a CUDA fault: the host must be able to tell
// the two apart (B200PROBE_ENOMEM -> "inconclusive", never "unhealthy").
#define B200_ALLOC_TRY(expr)                                                                                  \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ == cudaErrorMemoryAllocation) { cudaGetLastError(); b200::set_error("%s: out of device memory", #expr); return B200PROBE_ENOMEM; } \
        if (e__ != cudaSuccess) { b200::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); return b200::cuda_rc((int)e__); } \
    } while (0)

// ---- data pattern, shared host/device (header-only so the kernels inline it) -----------------
#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif
B200_HD uint32_t b200_pattern_word(uint64_t i, uint32_t seed) {
    return ((uint32_t)i * 2654435761u) ^ seed ^ (uint32_t)(i >> 32);
}
// splitmix-style 32-bit finaliser (GEMM operand generator)
B200_HD uint32_t b200_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// ---- GEMM operand class 1 (SURVEY.md §8d "GEMM probe": bf16 A,B ~ U(-1,1) from Philox seed 0xB200) ----------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123
// reference constants).  Counter-based: word j of block b is a pure function of (b, key), so the device fill, the
// host-side check and the oracle agree without sharing state.
B200_HD void b200_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// element e of matrix `which` (0 = A, 1 = B): word e%4 of Philox block (e/4, which, 0) under key (seed, 0);
// u = top 24 bits -> x = u * 2^-23 - 1 in [-1, 1) (exact in fp32) -> bf16 round-to-nearest-even.
B200_HD uint16_t b200_gemm_uniform_bits(uint64_t e, uint32_t seed, int which) {
    uint32_t r[4];
    const uint64_t blk = e >> 2;
    b200_philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)which, 0u, seed, 0u, r);
    const float x = (float)(r[e & 3] >> 8) * (1.0f / 8388608.0f) - 1.0f;
    uint32_t u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(x);
#else
    memcpy(&u, &x, 4);
#endif
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

// bf16 operand element e of matrix `which`: k/128 with k an integer in [-128,127], exactly
// representable in bf16 (8 significant bits), so the CPU oracle and the device agree on the bits.
B200_HD uint16_t b200_gemm_elem_bits(uint64_t e, uint32_t seed, int which) {
    uint32_t h = b200_mix32((uint32_t)e * 0x9E3779B1u ^ b200_mix32(seed + 0x51ED27u * (uint32_t)(which + 1)) ^ (uint32_t)(e >> 32));
    int k = (int)(h & 0xFF) - 128;                    // [-128,127]
    float f = (float)k * (1.0f / 128.0f);             // exact
    uint32_t u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    return (uint16_t)(u >> 16);                       // exact: low 16 mantissa bits are zero
}
