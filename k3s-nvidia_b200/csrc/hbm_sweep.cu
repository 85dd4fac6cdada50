// HBM bandwidth sweep kernels for sm_100a (no reference counterpart —
// the reference plugin's health path, /root/reference/README.md:116, moves no GPU bytes).
//
// Two families, same data contract (include/b200probe.h):
//   TMA ring  : every warp owns a ring of shared-memory stages.  global->shared moves are
//               cp.async.bulk (SASS UBLKCP) completing on an mbarrier; shared->global moves are
//               cp.async.bulk bulk_group stores.  The copy path never touches registers: one lane
//               per warp drives the TMA engine, so bytes in flight are bounded by shared memory
//               (up to ~200 KB/SM), not by registers or LSU queues.
//   direct    : LDG.128 / STG.128 grid-stride with 8 independent requests per thread.
//
// Roofline: HBM bound.  Algorithmic bytes per launch: read N, write N, copy 2N.
#include <cuda_runtime.h>
#include <pthread.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace {

using namespace b200ptx;

constexpr int kMaxWarps = 8;
constexpr int kMaxStages = 8;
constexpr int kModeVerify = 8;    // internal: READ + compare with the regenerated pattern (partials[2] bad words, [3] first bad word)

struct RingArgs {
    const uint8_t* src;
    uint8_t* dst;
    uint64_t bytes;            // multiple of 4
    uint32_t seed;
    uint32_t stage_bytes;      // multiple of 512
    uint32_t stages;           // 2..kMaxStages
    unsigned long long* partials;   // [0] sum64, [1] xor (read mode); verify adds [2] mismatching words, [3] lowest mismatching word
    // experiment knobs (environment only, never part of the ABI; 0 = the shipped behaviour):
    int load_policy, store_policy;  // L2 eviction policy of the bulk loads / stores: 0 evict_first, 1 evict_normal, 2 evict_last, 3 evict_unchanged
    int slab_map;                   // 1 = every worker owns one contiguous slab instead of every nworkers-th chunk
    int batch;                      // copy: 1 = fill ALL stages, then store ALL stages (per-warp read and write phases)
};

__device__ __forceinline__ void accum16(const uint4& v, unsigned long long& sum, uint32_t& x) {
    sum += (unsigned long long)v.x + v.y;
    sum += (unsigned long long)v.z + v.w;
    x ^= v.x ^ v.y ^ v.z ^ v.w;
}

__device__ __forceinline__ void warp_publish(unsigned long long sum, uint32_t x, unsigned long long* partials) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(partials, sum);
        atomicXor(partials + 1, (unsigned long long)x);
    }
}

// The <16-byte (word-granular) tail that bulk copies cannot carry; one thread of the grid.
__device__ __forceinline__ void tail_words(int mode, const RingArgs& a, uint64_t bulk_bytes) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long sum = 0;
    uint32_t x = 0;
    for (uint64_t off = bulk_bytes; off < a.bytes; off += 4) {
        if (mode == B200PROBE_HBM_READ || mode == kModeVerify) {
            uint32_t w = *reinterpret_cast<const uint32_t*>(a.src + off);
            sum += w; x ^= w;
            if (mode == kModeVerify && w != b200_pattern_word(off >> 2, a.seed)) {
                atomicAdd(a.partials + 2, 1ull);
                atomicMin(a.partials + 3, (unsigned long long)(off >> 2));
            }
        } else if (mode == B200PROBE_HBM_WRITE) {
            *reinterpret_cast<uint32_t*>(a.dst + off) = b200_pattern_word(off >> 2, a.seed);
        } else {
            *reinterpret_cast<uint32_t*>(a.dst + off) = *reinterpret_cast<const uint32_t*>(a.src + off);
        }
    }
    if ((mode == B200PROBE_HBM_READ || mode == kModeVerify) && bulk_bytes < a.bytes) {
        atomicAdd(a.partials, sum);
        atomicXor(a.partials + 1, (unsigned long long)x);
    }
}

// ------------------------------------------------------------------------------------------------
// TMA ring kernel.  One ring per warp; chunk c of the buffer goes to worker c % nworkers.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kMaxWarps * 32) hbm_ring_kernel(RingArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxWarps * kMaxStages];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const uint32_t S = a.stages, SB = a.stage_bytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kMaxWarps * kMaxStages; ++i) mbar_init(smem_u32(&full_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint64_t bulk_bytes = a.bytes & ~15ull;
    const uint64_t nchunks = (bulk_bytes + SB - 1) / SB;
    const uint64_t worker = (uint64_t)blockIdx.x * nwarps + warp;
    const uint64_t nworkers = (uint64_t)gridDim.x * nwarps;
    const uint64_t n_my = a.slab_map ? (worker * ((nchunks + nworkers - 1) / nworkers) < nchunks
                                            ? min((nchunks + nworkers - 1) / nworkers, nchunks - worker * ((nchunks + nworkers - 1) / nworkers)) : 0)
                                     : (worker < nchunks ? (nchunks - worker + nworkers - 1) / nworkers : 0);

    const uint32_t ring = smem_u32(smem) + warp * S * SB;
    uint8_t* ring_ptr = smem + (size_t)warp * S * SB;
    const uint32_t bar0 = smem_u32(&full_bar[warp * kMaxStages]);
    const uint64_t pol = policy_of(a.load_policy), pol_st = policy_of(a.store_policy);
    const uint64_t slab = (nchunks + nworkers - 1) / nworkers;     // slab_map: chunks [worker*slab, +slab)

    auto chunk_off = [&](uint64_t k) { return (a.slab_map ? worker * slab + k : worker + k * nworkers) * (uint64_t)SB; };
    auto chunk_len = [&](uint64_t k) { uint64_t o = chunk_off(k); return (uint32_t)min((uint64_t)SB, bulk_bytes - o); };
    // producer cursor (stage ps) and consumer cursor (stage cs, parity cph) advance without div/mod
    uint32_t ps = 0, cs = 0, cph = 0;
    uint64_t issued = 0;
    auto load_next = [&]() {
        const uint32_t len = chunk_len(issued);
        mbar_expect_tx(bar0 + ps * 8, len);
        bulk_g2s(ring + ps * SB, a.src + chunk_off(issued), len, bar0 + ps * 8, pol);
        ++issued;
        if (++ps == S) ps = 0;
    };
    auto advance = [&]() { if (++cs == S) { cs = 0; cph ^= 1; } };

    if (MODE == B200PROBE_HBM_COPY && a.batch) {
        if (lane == 0) {
            for (uint64_t k0 = 0; k0 < n_my; k0 += S) {           // read phase: S loads; write phase: S stores; repeat
                const uint64_t n = min((uint64_t)S, n_my - k0);
                for (uint64_t j = 0; j < n; ++j) load_next();
                for (uint64_t j = 0; j < n; ++j) {
                    mbar_wait(bar0 + cs * 8, cph);
                    bulk_s2g(a.dst + chunk_off(k0 + j), ring + cs * SB, chunk_len(k0 + j), pol_st);
                    advance();
                }
                bulk_commit();
                bulk_wait_read<0>();
            }
            bulk_wait_all();
        }
    } else if (MODE == B200PROBE_HBM_COPY) {
        if (lane == 0 && n_my > 0) {
            const uint64_t ahead = min((uint64_t)(S - 1), n_my);
            while (issued < ahead) load_next();
            for (uint64_t k = 0; k < n_my; ++k) {
                mbar_wait(bar0 + cs * 8, cph);
                bulk_s2g(a.dst + chunk_off(k), ring + cs * SB, chunk_len(k), pol_st);
                bulk_commit();
                if (issued < n_my) {
                    bulk_wait_read<1>();       // store k-1 has drained its stage, which the next load reuses
                    load_next();
                }
                advance();
            }
            bulk_wait_all();
        }
    } else if (MODE == B200PROBE_HBM_READ || MODE == kModeVerify) {
        unsigned long long sum = 0;
        uint32_t x = 0;
        if (lane == 0) {
            const uint64_t ahead = min((uint64_t)S, n_my);
            while (issued < ahead) load_next();
        }
        for (uint64_t k = 0; k < n_my; ++k) {
            mbar_wait(bar0 + cs * 8, cph);
            const uint4* st = reinterpret_cast<const uint4*>(ring_ptr + (size_t)cs * SB);
            const uint32_t nvec = chunk_len(k) >> 4;
            uint32_t i = lane;
            if (MODE == kModeVerify) {
                // word w of the buffer must be (u32)w * K ^ seed ^ (w >> 32); a chunk never straddles a 2^32-word
                // boundary (chunk offsets are multiples of the stage size), so the high part folds into one constant
                const uint64_t w0 = chunk_off(k) >> 2;
                const uint32_t c = a.seed ^ (uint32_t)(w0 >> 32);
                const uint32_t K = 2654435761u;
                const uint32_t e0 = (uint32_t)w0 * K;
                uint32_t diff = 0;
                auto cmp = [&](const uint4& v, uint32_t vi) {
                    const uint32_t e = e0 + vi * (4u * K);
                    diff |= (v.x ^ e ^ c) | (v.y ^ (e + K) ^ c) | (v.z ^ (e + 2u * K) ^ c) | (v.w ^ (e + 3u * K) ^ c);
                };
                for (; i + 96 < nvec; i += 128) {
                    uint4 v0 = st[i], v1 = st[i + 32], v2 = st[i + 64], v3 = st[i + 96];
                    accum16(v0, sum, x); accum16(v1, sum, x); accum16(v2, sum, x); accum16(v3, sum, x);
                    cmp(v0, i); cmp(v1, i + 32); cmp(v2, i + 64); cmp(v3, i + 96);
                }
                for (; i < nvec; i += 32) { uint4 v = st[i]; accum16(v, sum, x); cmp(v, i); }
                if (__any_sync(0xffffffffu, diff != 0)) {
                    // slow path, only for a stage that holds a fault: exact count and lowest index
                    unsigned long long bad = 0, first = ~0ull;
                    const uint32_t* wds = reinterpret_cast<const uint32_t*>(st);
                    for (uint32_t j = lane; j < nvec * 4; j += 32)
                        if (wds[j] != b200_pattern_word(w0 + j, a.seed)) { ++bad; first = min(first, (unsigned long long)(w0 + j)); }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        bad += __shfl_xor_sync(0xffffffffu, bad, o);
                        first = min(first, (unsigned long long)__shfl_xor_sync(0xffffffffu, first, o));
                    }
                    if (lane == 0 && bad) { atomicAdd(a.partials + 2, bad); atomicMin(a.partials + 3, first); }
                }
            } else {
                for (; i + 96 < nvec; i += 128) {
                    uint4 v0 = st[i], v1 = st[i + 32], v2 = st[i + 64], v3 = st[i + 96];
                    accum16(v0, sum, x); accum16(v1, sum, x); accum16(v2, sum, x); accum16(v3, sum, x);
                }
                for (; i < nvec; i += 32) accum16(st[i], sum, x);
            }
            __syncwarp();
            // lane 0 is the only producer: its `issued` counts k+S loads here, stage ps == cs
            if (lane == 0 && k + S < n_my) load_next();
            advance();
        }
        warp_publish(sum, x, a.partials);
    } else {   // WRITE: generate the pattern into the stage, then bulk-store it
        for (uint64_t k = 0; k < n_my; ++k) {
            if (k >= S) {
                if (lane == 0) bulk_wait_read_dyn((int)S - 1);   // store k-S no longer reads stage cs
                __syncwarp();
            }
            uint4* st = reinterpret_cast<uint4*>(ring_ptr + (size_t)cs * SB);
            const uint32_t nvec = chunk_len(k) >> 4;
            const uint64_t w0 = chunk_off(k) >> 2;
#pragma unroll 4
            for (uint32_t i = lane; i < nvec; i += 32) {
                uint64_t w = w0 + (uint64_t)i * 4;
                st[i] = make_uint4(b200_pattern_word(w, a.seed), b200_pattern_word(w + 1, a.seed),
                                   b200_pattern_word(w + 2, a.seed), b200_pattern_word(w + 3, a.seed));
            }
            fence_proxy_async_smem();          // generic-proxy writes -> visible to the async proxy
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(a.dst + chunk_off(k), ring + cs * SB, chunk_len(k), pol_st);
                bulk_commit();
            }
            advance();
        }
        if (lane == 0) bulk_wait_all();
    }
    tail_words(MODE, a, bulk_bytes);
}

// ------------------------------------------------------------------------------------------------
// Direct LDG.128 / STG.128 kernels: 8 independent 16-byte requests per thread per iteration.
// ------------------------------------------------------------------------------------------------
constexpr int kUnroll = 8;

template <int MODE>
__global__ void __launch_bounds__(512) hbm_direct_kernel(RingArgs a) {
    const uint64_t nvec = a.bytes >> 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(a.src);
    uint4* __restrict__ dst = reinterpret_cast<uint4*>(a.dst);
    unsigned long long sum = 0;
    uint32_t x = 0;

    uint64_t i = tid;
    for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
        uint4 v[kUnroll];
        if (MODE != B200PROBE_HBM_WRITE) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) v[j] = __ldcs(src + i + j * stride);
        } else {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                uint64_t w = (i + j * stride) * 4;
                v[j] = make_uint4(b200_pattern_word(w, a.seed), b200_pattern_word(w + 1, a.seed), b200_pattern_word(w + 2, a.seed),
                                  b200_pattern_word(w + 3, a.seed));
            }
        }
        if (MODE == B200PROBE_HBM_READ) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) accum16(v[j], sum, x);
        } else {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) __stcs(dst + i + j * stride, v[j]);
        }
    }
    for (; i < nvec; i += stride) {
        uint4 v;
        if (MODE != B200PROBE_HBM_WRITE) v = __ldcs(src + i);
        else {
            uint64_t w = i * 4;
            v = make_uint4(b200_pattern_word(w, a.seed), b200_pattern_word(w + 1, a.seed), b200_pattern_word(w + 2, a.seed),
                           b200_pattern_word(w + 3, a.seed));
        }
        if (MODE == B200PROBE_HBM_READ) accum16(v, sum, x);
        else __stcs(dst + i, v);
    }
    if (MODE == B200PROBE_HBM_READ) warp_publish(sum, x, a.partials);
    tail_words(MODE, a, a.bytes & ~15ull);
}

// ------------------------------------------------------------------------------------------------
// Verification kernel: one streaming pass that (a) checksums the buffer (the data result compared
// with the oracle) and (b) regenerates the closed-form pattern and counts words that differ
// (the probe's own health verdict: a flipped bit anywhere in the swept range shows up here).
// partials: [0] sum64, [1] xor32, [2] mismatching words, [3] index of the lowest mismatch.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) hbm_verify_kernel(const uint8_t* __restrict__ buf, uint64_t bytes, uint32_t seed,
                                                         unsigned long long* partials) {
    const uint64_t nvec = bytes >> 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(buf);
    unsigned long long sum = 0, bad = 0, first = ~0ull;
    uint32_t x = 0;
    auto check = [&](const uint4& v, uint64_t i) {
        accum16(v, sum, x);
        const uint64_t w = i * 4;
        const uint32_t d = (v.x ^ b200_pattern_word(w, seed)) | (v.y ^ b200_pattern_word(w + 1, seed)) |
                           (v.z ^ b200_pattern_word(w + 2, seed)) | (v.w ^ b200_pattern_word(w + 3, seed));
        if (d) {
            const bool bx = v.x != b200_pattern_word(w, seed), by = v.y != b200_pattern_word(w + 1, seed), bz = v.z != b200_pattern_word(w + 2, seed),
                       bw = v.w != b200_pattern_word(w + 3, seed);
            bad += bx + by + bz + bw;
            first = min(first, (unsigned long long)(w + (bx ? 0 : by ? 1 : bz ? 2 : 3)));     // exact word, not the vector it sits in
        }
    };
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        uint4 v0 = __ldcs(src + i), v1 = __ldcs(src + i + stride), v2 = __ldcs(src + i + 2 * stride), v3 = __ldcs(src + i + 3 * stride);
        check(v0, i); check(v1, i + stride); check(v2, i + 2 * stride); check(v3, i + 3 * stride);
    }
    for (; i < nvec; i += stride) check(__ldcs(src + i), i);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint64_t off = bytes & ~15ull; off < bytes; off += 4) {
            uint32_t w = *reinterpret_cast<const uint32_t*>(buf + off);
            sum += w; x ^= w;
            if (w != b200_pattern_word(off >> 2, seed)) { ++bad; first = min(first, (unsigned long long)(off >> 2)); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        first = min(first, (unsigned long long)__shfl_xor_sync(0xffffffffu, first, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(partials, sum);
        atomicXor(partials + 1, (unsigned long long)x);
        if (bad) { atomicAdd(partials + 2, bad); atomicMin(partials + 3, first); }
    }
}

// ------------------------------------------------------------------------------------------------
// Launch plumbing
// ------------------------------------------------------------------------------------------------
struct Tuning {
    int variant, stage_bytes, stages, warps, ctas_per_sm;
};

// Defaults per mode from the 1 GiB tuning sweep on B200 (profiles/hbm_tune_r01_1GiB.txt):
//   copy  8 KiB x 4 stages x 2 warps, 1 CTA/SM  -> 6515 GB/s (torch copy_ 6400)
//   read  8 KiB x 3 stages x 4 warps            -> 6890 GB/s
//   write 8 KiB x 6 stages x 4 warps            -> 7024 GB/s
// More bytes in flight than ~64-190 KB/SM makes copy slower (DRAM read/write turnarounds), not faster.
Tuning resolve_tuning(const b200probe_hbm_cfg_t* c, int mode) {
    Tuning t;
    t.variant = c ? c->variant : B200PROBE_VARIANT_TMA;
    const int d_stages = mode == B200PROBE_HBM_COPY ? 4 : (mode == B200PROBE_HBM_READ || mode == kModeVerify) ? 3 : 6;
    const int d_warps = mode == B200PROBE_HBM_COPY ? 2 : 4;
    t.stage_bytes = (c && c->stage_bytes) ? c->stage_bytes : 8192;
    t.stages = (c && c->stages) ? c->stages : d_stages;
    t.warps = (c && c->warps_per_cta) ? c->warps_per_cta : d_warps;
    t.ctas_per_sm = (c && c->ctas_per_sm) ? c->ctas_per_sm : (t.variant == B200PROBE_VARIANT_TMA ? 1 : (mode == B200PROBE_HBM_READ ? 3 : 8));
    return t;
}

int check_args(const void* p0, const void* p1, uint64_t bytes) {
    if (bytes & 3) { b200::set_error("bytes=%llu is not a multiple of 4", (unsigned long long)bytes); return B200PROBE_EINVAL; }
    if (((uintptr_t)p0 | (uintptr_t)p1) & 15) { b200::set_error("device pointers must be 16-byte aligned"); return B200PROBE_EINVAL; }
    return 0;
}

int env_int(const char* name) {
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
}

template <int MODE>
int launch_mode(int ordinal, RingArgs a, const b200probe_hbm_cfg_t* cfg, cudaStream_t stream) {
    b200::DevProps props;
    int rc = b200::device_props(ordinal, &props);
    if (rc) return rc;
    if (a.bytes == 0) return 0;
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    Tuning t = resolve_tuning(cfg, MODE);
    if (t.variant == B200PROBE_VARIANT_TMA) {
        if (t.stages < 2 || t.stages > kMaxStages || t.warps < 1 || t.warps > kMaxWarps || t.stage_bytes < 512 || (t.stage_bytes & 511) ||
            t.ctas_per_sm < 1) {
            b200::set_error("bad TMA tuning: stage_bytes=%d stages=%d warps=%d ctas_per_sm=%d", t.stage_bytes, t.stages, t.warps, t.ctas_per_sm);
            return B200PROBE_EINVAL;
        }
        size_t smem = (size_t)t.warps * t.stages * t.stage_bytes;
        if (smem + 1024 > (size_t)props.smem_optin) {
            b200::set_error("ring needs %zu B shared memory, device allows %d", smem, props.smem_optin);
            return B200PROBE_EINVAL;
        }
        a.stage_bytes = (uint32_t)t.stage_bytes;
        a.stages = (uint32_t)t.stages;
        a.load_policy = env_int("B200PROBE_HBM_LOAD_POLICY"); a.store_policy = env_int("B200PROBE_HBM_STORE_POLICY");      // read per launch:
        a.slab_map = env_int("B200PROBE_HBM_SLAB_MAP"); a.batch = env_int("B200PROBE_HBM_COPY_BATCH");                     // tools flip them in-process
        B200_CUDA_TRY(cudaFuncSetAttribute(hbm_ring_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = props.sms * t.ctas_per_sm;
        hbm_ring_kernel<MODE><<<grid, t.warps * 32, smem, stream>>>(a);
    } else if (t.variant == B200PROBE_VARIANT_DIRECT) {
        int grid = props.sms * t.ctas_per_sm;
        hbm_direct_kernel<MODE><<<grid, 512, 0, stream>>>(a);
    } else {
        b200::set_error("unknown variant %d", t.variant);
        return B200PROBE_EINVAL;
    }
    B200_CUDA_TRY(cudaGetLastError());
    return 0;
}

double median(std::vector<float> v) {
    std::sort(v.begin(), v.end());
    size_t n = v.size();
    return n & 1 ? v[n / 2] : 0.5 * (v[n / 2 - 1] + v[n / 2]);
}

// Per-device resident arena: the probe runs periodically inside the plugin, so its buffers,
// stream and events stay allocated between calls (cudaMalloc/cudaFree of 2 GiB costs more than
// the sweep itself).  Released by b200probe_hbm_release / b200probe_shutdown.
struct Arena {
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    uint8_t *src = nullptr, *dst = nullptr, *flush = nullptr;
    uint64_t cap = 0, flush_cap = 0;
    unsigned long long* partials = nullptr;     // 4 x u64
    cudaStream_t stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;   // host round trip: H2D and D2H legs overlap the kernel leg
    std::vector<cudaEvent_t> chunk_ev;              // 2 per chunk in flight (landed on device / kernel done)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    uint64_t src_bytes = 0;                     // prefix of src currently holding pattern(src_seed)
    uint32_t src_seed = 0;
    bool inited = false;                        // stream, events and partials all exist
};
Arena g_arena[B200PROBE_MAX_DEVICES];

struct ArenaLock {
    Arena& a;
    explicit ArenaLock(Arena& x) : a(x) { pthread_mutex_lock(&a.mu); }
    ~ArenaLock() { pthread_mutex_unlock(&a.mu); }
};

void arena_free(Arena& a) {
    if (a.src) cudaFree(a.src);
    if (a.dst) cudaFree(a.dst);
    if (a.flush) cudaFree(a.flush);
    if (a.partials) cudaFree(a.partials);
    if (a.e0) cudaEventDestroy(a.e0);
    if (a.e1) cudaEventDestroy(a.e1);
    if (a.stream) cudaStreamDestroy(a.stream);
    if (a.s_in) cudaStreamDestroy(a.s_in);
    if (a.s_out) cudaStreamDestroy(a.s_out);
    for (cudaEvent_t e : a.chunk_ev) cudaEventDestroy(e);
    a.chunk_ev.clear();
    a.s_in = a.s_out = nullptr;
    a.src = a.dst = a.flush = nullptr;
    a.partials = nullptr; a.e0 = a.e1 = nullptr; a.stream = nullptr;
    a.cap = a.flush_cap = a.src_bytes = 0;
    a.inited = false;
}

int arena_reserve_inner(Arena& a, uint64_t bytes, uint64_t flush_bytes) {
    if (!a.inited) {
        B200_CUDA_TRY(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
        B200_CUDA_TRY(cudaEventCreate(&a.e0));
        B200_CUDA_TRY(cudaEventCreate(&a.e1));
        B200_ALLOC_TRY(cudaMalloc(&a.partials, 32));
        a.inited = true;                    // only now: a half-built arena is torn down by the caller, never reused
    }
    if (bytes > a.cap) {
        if (a.src) cudaFree(a.src);
        if (a.dst) cudaFree(a.dst);
        a.src = a.dst = nullptr; a.cap = 0; a.src_bytes = 0;
        B200_ALLOC_TRY(cudaMalloc(&a.src, bytes));
        B200_ALLOC_TRY(cudaMalloc(&a.dst, bytes));
        a.cap = bytes;
    }
    if (flush_bytes > a.flush_cap) {
        if (a.flush) cudaFree(a.flush);
        a.flush = nullptr; a.flush_cap = 0;
        B200_ALLOC_TRY(cudaMalloc(&a.flush, flush_bytes));
        a.flush_cap = flush_bytes;
    }
    return 0;
}

// Any failure (out of memory included) leaves NO arena behind: the next call starts from scratch instead of finding a stream
// without its events, or a source buffer without its destination.
int arena_reserve(Arena& a, uint64_t bytes, uint64_t flush_bytes) {
    const int rc = arena_reserve_inner(a, bytes, flush_bytes);
    if (rc) arena_free(a);
    return rc;
}

// Verification launch: the TMA-ring read kernel with the pattern compare folded in (same staging and
// tuning as the read sweep, so the verdict pass runs at read bandwidth); B200PROBE_VERIFY_DIRECT=1 selects
// the LDG grid-stride kernel instead (A/B checks).
int launch_verify(int ordinal, const uint8_t* buf, uint64_t bytes, uint32_t seed, unsigned long long* partials, cudaStream_t stream, int sms) {
    static const bool direct = [] { const char* e = getenv("B200PROBE_VERIFY_DIRECT"); return e && *e == '1'; }();
    if (direct || (((uintptr_t)buf) & 15)) {
        hbm_verify_kernel<<<sms * 4, 512, 0, stream>>>(buf, bytes, seed, partials);
        B200_CUDA_TRY(cudaGetLastError());
        return 0;
    }
    RingArgs a{buf, nullptr, bytes, seed, 0, 0, partials, 0, 0, 0, 0};
    return launch_mode<kModeVerify>(ordinal, a, nullptr, stream);
}

struct VerifyOut { uint64_t sum; uint32_t x; uint64_t bad, first; };

int verify_pass(int ordinal, const uint8_t* buf, uint64_t bytes, uint32_t seed, Arena& a, int sms, VerifyOut* v) {
    static const unsigned long long init[4] = {0, 0, 0, ~0ull};
    B200_CUDA_TRY(cudaMemcpyAsync(a.partials, init, 32, cudaMemcpyHostToDevice, a.stream));
    if (bytes) {
        int rc = launch_verify(ordinal, buf, bytes, seed, a.partials, a.stream, sms);
        if (rc) return rc;
    }
    unsigned long long h[4];
    B200_CUDA_TRY(cudaMemcpyAsync(h, a.partials, 32, cudaMemcpyDeviceToHost, a.stream));
    B200_CUDA_TRY(cudaStreamSynchronize(a.stream));
    v->sum = h[0]; v->x = (uint32_t)h[1]; v->bad = h[2]; v->first = h[3];
    (void)ordinal;
    return 0;
}

}  // namespace

namespace b200 {
int verify_pattern(int ordinal, const void* buf, uint64_t bytes, uint32_t seed, unsigned long long* d_partials, void* stream, VerifyResult* out) {
    DevProps props;
    int rc = device_props(ordinal, &props);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    static const unsigned long long init[4] = {0, 0, 0, ~0ull};
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    B200_CUDA_TRY(cudaMemcpyAsync(d_partials, init, 32, cudaMemcpyHostToDevice, st));
    if (bytes) {
        rc = launch_verify(ordinal, (const uint8_t*)buf, bytes, seed, d_partials, st, props.sms);
        if (rc) return rc;
    }
    unsigned long long h[4];
    B200_CUDA_TRY(cudaMemcpyAsync(h, d_partials, 32, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    out->sum = h[0]; out->x = (uint32_t)h[1]; out->bad = h[2]; out->first = h[3];
    return 0;
}
}  // namespace b200

extern "C" {

int b200probe_hbm_fill(int ordinal, void* dst, uint64_t bytes, uint32_t seed, const b200probe_hbm_cfg_t* cfg, void* stream) {
    int rc = check_args(dst, nullptr, bytes);
    if (rc) return rc;
    RingArgs a{nullptr, (uint8_t*)dst, bytes, seed, 0, 0, nullptr, 0, 0, 0, 0};
    return launch_mode<B200PROBE_HBM_WRITE>(ordinal, a, cfg, (cudaStream_t)stream);
}

int b200probe_hbm_copy(int ordinal, const void* src, void* dst, uint64_t bytes, const b200probe_hbm_cfg_t* cfg, void* stream) {
    int rc = check_args(src, dst, bytes);
    if (rc) return rc;
    RingArgs a{(const uint8_t*)src, (uint8_t*)dst, bytes, 0, 0, 0, nullptr, 0, 0, 0, 0};
    return launch_mode<B200PROBE_HBM_COPY>(ordinal, a, cfg, (cudaStream_t)stream);
}

int b200probe_hbm_read(int ordinal, const void* src, uint64_t bytes, uint64_t* partials, const b200probe_hbm_cfg_t* cfg, void* stream) {
    int rc = check_args(src, partials, bytes);
    if (rc) return rc;
    if (!partials) return B200PROBE_EINVAL;
    RingArgs a{(const uint8_t*)src, nullptr, bytes, 0, 0, 0, (unsigned long long*)partials, 0, 0, 0, 0};
    return launch_mode<B200PROBE_HBM_READ>(ordinal, a, cfg, (cudaStream_t)stream);
}

int b200probe_hbm_verify(int ordinal, const void* buf, uint64_t bytes, uint32_t seed, uint64_t* sum64, uint32_t* xor32,
                         uint64_t* bad_words, uint64_t* first_bad_word) {
    int rc = check_args(buf, nullptr, bytes);
    if (rc) return rc;
    b200::DevProps props;
    rc = b200::device_props(ordinal, &props);
    if (rc) return rc;
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    Arena& a = g_arena[ordinal];
    ArenaLock lock(a);
    rc = arena_reserve(a, 0, 0);
    if (rc) return rc;
    VerifyOut v;
    rc = verify_pass(ordinal, (const uint8_t*)buf, bytes, seed, a, props.sms, &v);
    if (rc) return rc;
    if (sum64) *sum64 = v.sum;
    if (xor32) *xor32 = v.x;
    if (bad_words) *bad_words = v.bad;
    if (first_bad_word) *first_bad_word = v.first;
    return 0;
}

int b200probe_host_alloc(uint64_t bytes, void** ptr) {
    if (!ptr || !bytes) return B200PROBE_EINVAL;
    b200::DevProps props;
    int rc = b200::device_props(0, &props);          // CUDA must be usable: pinned memory belongs to the driver
    if (rc) return rc;
    B200_CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return 0;
}

int b200probe_host_free(void* ptr) {
    if (!ptr) return 0;
    B200_CUDA_TRY(cudaFreeHost(ptr));
    return 0;
}

// Host round trip as a three-leg pipeline over chunks: H2D (s_in) -> copy kernel + checksum (stream)
// -> D2H (s_out).  With pinned host buffers (b200probe_host_alloc) the two PCIe directions run
// concurrently and the kernel leg hides under them; pageable buffers still work (the driver stages them).
int b200probe_hbm_copy_host(int ordinal, const void* src_host, void* dst_host, uint64_t bytes, uint64_t* sum64, uint32_t* xor32) {
    if ((bytes & 3) || (bytes && (!src_host || !dst_host))) { b200::set_error("copy_host: bad arguments"); return B200PROBE_EINVAL; }
    b200::DevProps props;
    int rc = b200::device_props(ordinal, &props);
    if (rc) return rc;
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    Arena& a = g_arena[ordinal];
    ArenaLock lock(a);
    rc = arena_reserve(a, std::max<uint64_t>(bytes, 16), 0);
    if (rc) return rc;
    if (!a.s_in) {
        B200_CUDA_TRY(cudaStreamCreateWithFlags(&a.s_in, cudaStreamNonBlocking));
        B200_CUDA_TRY(cudaStreamCreateWithFlags(&a.s_out, cudaStreamNonBlocking));
    }
    a.src_bytes = 0;   // src no longer holds the pattern
    uint64_t chunk = 8ull << 20;
    if (const char* e = getenv("B200PROBE_HOST_CHUNK_BYTES")) { long long v = atoll(e); if (v >= 65536) chunk = (uint64_t)v & ~15ull; }
    while ((bytes + chunk - 1) / chunk > 64) chunk <<= 1;
    const size_t nchunks = (size_t)((bytes + chunk - 1) / chunk);
    while (a.chunk_ev.size() < 2 * nchunks) {
        cudaEvent_t e;
        B200_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        a.chunk_ev.push_back(e);
    }
    B200_CUDA_TRY(cudaMemsetAsync(a.partials, 0, 32, a.stream));
    for (size_t c = 0; c < nchunks; ++c) {
        const uint64_t off = c * chunk, len = std::min<uint64_t>(chunk, bytes - off);
        B200_CUDA_TRY(cudaMemcpyAsync(a.src + off, (const uint8_t*)src_host + off, len, cudaMemcpyHostToDevice, a.s_in));
        B200_CUDA_TRY(cudaEventRecord(a.chunk_ev[2 * c], a.s_in));
        B200_CUDA_TRY(cudaStreamWaitEvent(a.stream, a.chunk_ev[2 * c], 0));
        rc = b200probe_hbm_copy(ordinal, a.src + off, a.dst + off, len, nullptr, a.stream);
        if (rc) return rc;
        rc = b200probe_hbm_read(ordinal, a.dst + off, len, (uint64_t*)a.partials, nullptr, a.stream);   // checksum of what landed
        if (rc) return rc;
        B200_CUDA_TRY(cudaEventRecord(a.chunk_ev[2 * c + 1], a.stream));
        B200_CUDA_TRY(cudaStreamWaitEvent(a.s_out, a.chunk_ev[2 * c + 1], 0));
        B200_CUDA_TRY(cudaMemcpyAsync((uint8_t*)dst_host + off, a.dst + off, len, cudaMemcpyDeviceToHost, a.s_out));
    }
    unsigned long long h[4] = {0, 0, 0, 0};
    B200_CUDA_TRY(cudaMemcpyAsync(h, a.partials, 32, cudaMemcpyDeviceToHost, a.stream));
    B200_CUDA_TRY(cudaStreamSynchronize(a.stream));
    B200_CUDA_TRY(cudaStreamSynchronize(a.s_out));
    if (sum64) *sum64 = h[0];
    if (xor32) *xor32 = (uint32_t)h[1];
    return 0;
}

int b200probe_hbm_release(int ordinal) {
    if (ordinal < 0 || ordinal >= B200PROBE_MAX_DEVICES) return B200PROBE_ERANGE;
    Arena& a = g_arena[ordinal];
    ArenaLock lock(a);
    if (a.stream || a.src) { cudaSetDevice(ordinal); arena_free(a); }
    return 0;
}

int b200probe_hbm_sweep(int idx, const b200probe_hbm_cfg_t* cfg_in, b200probe_hbm_result_t* out, int cap, int* n_out) {
    if (!out || !n_out || cap <= 0) return B200PROBE_EINVAL;
    int ordinal = -1;
    int rc = b200::cuda_ordinal_of(idx, &ordinal);
    if (rc) return rc;
    b200::DevProps props;
    rc = b200::device_props(ordinal, &props);
    if (rc) return rc;

    b200probe_hbm_cfg_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    if (cfg_in) cfg = *cfg_in;
    if (!cfg.min_bytes) cfg.min_bytes = 1ull << 20;
    if (!cfg.max_bytes) cfg.max_bytes = 1ull << 30;
    if (!cfg.modes) cfg.modes = B200PROBE_HBM_READ | B200PROBE_HBM_WRITE | B200PROBE_HBM_COPY;
    if (!cfg.warmup && !cfg.reps) { cfg.warmup = 3; cfg.reps = 20; }
    if (cfg.reps < 1) cfg.reps = 1;
    if (!cfg.seed) cfg.seed = 0xB200u;
    if (cfg.min_bytes > cfg.max_bytes || (cfg.min_bytes & (cfg.min_bytes - 1)) || (cfg.max_bytes & (cfg.max_bytes - 1)) || cfg.min_bytes < 16) {
        b200::set_error("sweep sizes must be powers of two, min<=max, min>=16");
        return B200PROBE_EINVAL;
    }

    B200_CUDA_TRY(cudaSetDevice(ordinal));
    Arena& a = g_arena[ordinal];
    ArenaLock lock(a);
    const size_t flush_bytes = cfg.flush_l2 ? (size_t)props.l2_bytes * 2 : 0;
    rc = arena_reserve(a, cfg.max_bytes, flush_bytes);
    if (rc) return rc;

    // source pattern (also the expected content of every mode's result); kept across calls
    if (a.src_bytes < cfg.max_bytes || a.src_seed != cfg.seed) {
        rc = b200probe_hbm_fill(ordinal, a.src, cfg.max_bytes, cfg.seed, &cfg, a.stream);
        if (rc) return rc;
        a.src_bytes = cfg.max_bytes;
        a.src_seed = cfg.seed;
    }

    int n = 0;
    const int mode_list[3] = {B200PROBE_HBM_READ, B200PROBE_HBM_WRITE, B200PROBE_HBM_COPY};
    std::vector<float> ms((size_t)cfg.reps);
    for (uint64_t bytes = cfg.min_bytes; bytes <= cfg.max_bytes; bytes <<= 1) {
        for (int mi = 0; mi < 3; ++mi) {
            const int mode = mode_list[mi];
            if (!(cfg.modes & mode)) continue;
            if (n >= cap) { b200::set_error("result buffer too small (cap=%d)", cap); return B200PROBE_ERANGE; }
            b200probe_hbm_result_t& r = out[n];
            memset(&r, 0, sizeof(r));
            r.bytes = bytes; r.mode = mode; r.variant = cfg.variant; r.verified = -1;
            const uint64_t footprint = mode == B200PROBE_HBM_COPY ? 2 * bytes : bytes;
            r.cache_resident = footprint <= (uint64_t)props.l2_bytes;
            if (mode != B200PROBE_HBM_READ && cfg.verify) B200_CUDA_TRY(cudaMemsetAsync(a.dst, 0, bytes, a.stream));
            B200_CUDA_TRY(cudaMemsetAsync(a.partials, 0, 32, a.stream));
            // One timed rep = `inner` back-to-back launches between two events (about a millisecond of work), so
            // the figure is the kernel's rate, not launch latency plus an event round trip; with an L2 flush
            // between reps the launches are timed one at a time.
            const double alg_bytes = (double)(mode == B200PROBE_HBM_COPY ? 2 * bytes : bytes);
            const int inner = cfg.flush_l2 ? 1 : cfg.launches_per_rep > 0 ? std::min(cfg.launches_per_rep, 1024)
                                                 : (int)std::max(1.0, std::min(64.0, std::floor(7.0e9 / alg_bytes + 0.5)));
            for (int it = -cfg.warmup; it < cfg.reps; ++it) {
                if (cfg.flush_l2) B200_CUDA_TRY(cudaMemsetAsync(a.flush, it & 0xff, flush_bytes, a.stream));
                if (it >= 0) B200_CUDA_TRY(cudaEventRecord(a.e0, a.stream));
                for (int k = 0; k < (it >= 0 ? inner : 1); ++k) {
                    if (mode == B200PROBE_HBM_READ) rc = b200probe_hbm_read(ordinal, a.src, bytes, (uint64_t*)a.partials, &cfg, a.stream);
                    else if (mode == B200PROBE_HBM_WRITE) rc = b200probe_hbm_fill(ordinal, a.dst, bytes, cfg.seed, &cfg, a.stream);
                    else rc = b200probe_hbm_copy(ordinal, a.src, a.dst, bytes, &cfg, a.stream);
                    if (rc) return rc;
                }
                if (it >= 0) {
                    B200_CUDA_TRY(cudaEventRecord(a.e1, a.stream));
                    B200_CUDA_TRY(cudaEventSynchronize(a.e1));
                    B200_CUDA_TRY(cudaEventElapsedTime(&ms[it], a.e0, a.e1));
                    ms[it] /= (float)inner;
                }
            }
            r.ms_median = median(ms);
            r.ms_best = *std::min_element(ms.begin(), ms.end());
            const double alg = (double)(mode == B200PROBE_HBM_COPY ? 2 * bytes : bytes);
            r.gbs_median = alg / (r.ms_median * 1e-3) / 1e9;
            r.gbs_best = alg / (r.ms_best * 1e-3) / 1e9;
            if (cfg.verify) {
                // one clean pass over the buffer the mode produced (write/copy) or consumed (read):
                // checksum = data result, mismatch count vs the regenerated pattern = verdict
                VerifyOut v;
                rc = verify_pass(ordinal, mode == B200PROBE_HBM_READ ? a.src : a.dst, bytes, cfg.seed, a, props.sms, &v);
                if (rc) return rc;
                r.sum64 = v.sum; r.xor32 = v.x;
                r.verified = v.bad == 0 ? 1 : 0;
                if (v.bad) {
                    b200::set_error("HBM sweep: %llu words differ from the pattern at %llu bytes mode %d (first bad word %llu)",
                                    (unsigned long long)v.bad, (unsigned long long)bytes, mode, (unsigned long long)v.first);
                    *n_out = n + 1;
                    return B200PROBE_EMISMATCH;
                }
            } else {
                B200_CUDA_TRY(cudaStreamSynchronize(a.stream));
            }
            ++n;
        }
    }
    *n_out = n;
    return 0;
}

}  // extern "C"
