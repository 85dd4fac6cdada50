// NVLink 5 all-to-all probe (no reference counterpart).
//
// Peer-memory kernels: every GPU pushes its (G-1) chunks straight into its peers' receive windows
// with 16-byte stores on peer-mapped pointers (NVSwitch routes them; no staging copy, no NCCL).
// The payload is either generated in registers (pure egress test: no local HBM read competes
// with the links) or read from a resident send buffer.  NCCL grouped send/recv is kept beside it
// as the library leg for contrast (dlopen'ed; mode B200PROBE_A2A_NCCL).
//
// Roofline: NVLink bound.  Algorithmic bytes per GPU per direction: (G-1)*S payload bytes.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace {

constexpr int kMaxWorld = 16;

struct PeerTable {
    uint8_t* win[kMaxWorld];   // base of every rank's receive window, peer-mapped (win[rank] is local)
};

B200_HD uint32_t chunk_seed(uint32_t seed, int src, int dst) { return seed ^ b200_mix32((uint32_t)(src * 251 + dst * 7 + 1)); }

// grid.x = world * ctas_per_peer; CTA b serves destination b / ctas_per_peer.
// only_dst >= 0 restricts the push to one destination (pairwise-isolated pass); -1 = all slots
// including the local one; -2 = all peers but not the local slot (pure NVLink traffic).
template <bool FROM_BUF>
__global__ void __launch_bounds__(512) a2a_push_kernel(PeerTable peers, const uint8_t* __restrict__ sendbuf, int rank, int world,
                                                       uint64_t S, uint32_t seed, int ctas_per_peer, int only_dst) {
    const int dst = blockIdx.x / ctas_per_peer;
    if (dst >= world) return;
    if (only_dst >= 0 && dst != only_dst) return;
    if (only_dst == -2 && dst == rank) return;      // -2: every peer, skip the local slot
    const int sub = blockIdx.x % ctas_per_peer;
    uint4* __restrict__ out = reinterpret_cast<uint4*>(peers.win[dst] + (uint64_t)rank * S);
    const uint4* __restrict__ in = reinterpret_cast<const uint4*>(sendbuf + (uint64_t)dst * S);
    const uint32_t cs = chunk_seed(seed, rank, dst);
    const uint64_t nvec = S >> 4;
    const uint64_t stride = (uint64_t)ctas_per_peer * blockDim.x;
    uint64_t i = (uint64_t)sub * blockDim.x + threadIdx.x;
    constexpr int U = 8;
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (FROM_BUF) v[j] = __ldcs(in + i + j * stride);
            else {
                uint64_t w = (i + j * stride) * 4;
                v[j] = make_uint4(b200_pattern_word(w, cs), b200_pattern_word(w + 1, cs), b200_pattern_word(w + 2, cs), b200_pattern_word(w + 3, cs));
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) out[i + j * stride] = v[j];
    }
    for (; i < nvec; i += stride) {
        uint4 v;
        if (FROM_BUF) v = __ldcs(in + i);
        else {
            uint64_t w = i * 4;
            v = make_uint4(b200_pattern_word(w, cs), b200_pattern_word(w + 1, cs), b200_pattern_word(w + 2, cs), b200_pattern_word(w + 3, cs));
        }
        out[i] = v;
    }
}


// TMA variant: each warp owns a ring of shared-memory stages, generates (or bulk-loads) a chunk
// into a stage and bulk-stores it (cp.async.bulk shared->global, SASS UBLKCP) to the PEER window:
// the copy engine of the SM streams whole 8-32 KiB bursts into NVLink instead of 16-byte stores.
constexpr int kRingWarps = 4, kRingMaxStages = 8;
template <bool FROM_BUF>
__global__ void __launch_bounds__(kRingWarps * 32) a2a_ring_push_kernel(PeerTable peers, const uint8_t* __restrict__ sendbuf, int rank, int world,
                                                                         uint64_t S, uint32_t seed, int ctas_per_peer, int only_dst,
                                                                         uint32_t SB, uint32_t NS) {
    using namespace b200ptx;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kRingWarps * kRingMaxStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kRingWarps * kRingMaxStages; ++i) mbar_init(smem_u32(&full_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int dst = blockIdx.x / ctas_per_peer;
    if (dst >= world) return;
    if (only_dst >= 0 && dst != only_dst) return;
    if (only_dst == -2 && dst == rank) return;
    const int sub = blockIdx.x % ctas_per_peer;
    uint8_t* out = peers.win[dst] + (uint64_t)rank * S;
    const uint8_t* in = sendbuf + (uint64_t)dst * S;
    const uint32_t cs = chunk_seed(seed, rank, dst);
    const uint64_t nchunks = (S + SB - 1) / SB;
    const uint64_t worker = (uint64_t)sub * kRingWarps + warp, nworkers = (uint64_t)ctas_per_peer * kRingWarps;
    const uint64_t n_my = worker < nchunks ? (nchunks - worker + nworkers - 1) / nworkers : 0;
    const uint32_t ring = smem_u32(smem) + warp * NS * SB;
    uint8_t* ring_ptr = smem + (size_t)warp * NS * SB;
    const uint32_t bar0 = smem_u32(&full_bar[warp * kRingMaxStages]);
    const uint64_t pol = policy_evict_first();
    auto off_of = [&](uint64_t k) { return (worker + k * nworkers) * (uint64_t)SB; };
    auto len_of = [&](uint64_t k) { return (uint32_t)min((uint64_t)SB, S - off_of(k)); };
    uint32_t cs_stage = 0, cph = 0;
    if (FROM_BUF) {
        if (lane == 0 && n_my > 0) {     // bulk-load local chunk, bulk-store to the peer; registers untouched
            uint64_t issued = 0;
            uint32_t ps = 0;
            auto load_next = [&]() {
                const uint32_t len = len_of(issued);
                mbar_expect_tx(bar0 + ps * 8, len);
                bulk_g2s(ring + ps * SB, in + off_of(issued), len, bar0 + ps * 8, pol);
                ++issued;
                if (++ps == NS) ps = 0;
            };
            const uint64_t ahead = min((uint64_t)(NS - 1), n_my);
            while (issued < ahead) load_next();
            for (uint64_t k = 0; k < n_my; ++k) {
                mbar_wait(bar0 + cs_stage * 8, cph);
                bulk_s2g(out + off_of(k), ring + cs_stage * SB, len_of(k), pol);
                bulk_commit();
                if (issued < n_my) { bulk_wait_read<1>(); load_next(); }
                if (++cs_stage == NS) { cs_stage = 0; cph ^= 1; }
            }
            bulk_wait_all();
        }
    } else {
        for (uint64_t k = 0; k < n_my; ++k) {
            if (k >= NS) {
                if (lane == 0) bulk_wait_read_dyn((int)NS - 1);
                __syncwarp();
            }
            uint4* st = reinterpret_cast<uint4*>(ring_ptr + (size_t)cs_stage * SB);
            const uint32_t nvec = len_of(k) >> 4;
            const uint64_t w0 = off_of(k) >> 2;
#pragma unroll 4
            for (uint32_t i = lane; i < nvec; i += 32) {
                const uint64_t w = w0 + (uint64_t)i * 4;
                st[i] = make_uint4(b200_pattern_word(w, cs), b200_pattern_word(w + 1, cs), b200_pattern_word(w + 2, cs), b200_pattern_word(w + 3, cs));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_s2g(out + off_of(k), ring + cs_stage * SB, len_of(k), pol); bulk_commit(); }
            if (++cs_stage == NS) cs_stage = 0;
        }
        if (lane == 0) bulk_wait_all();
    }
}

int default_ctas_per_peer(int sms, int world) { return std::max(1, (sms * 4 + world - 1) / world); }

int launch_push(int ordinal, int rank, int world, void* const* windows, const void* sendbuf, uint64_t S, uint32_t seed, int ctas_per_peer,
                int only_dst, int variant, cudaStream_t stream) {
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || !windows) { b200::set_error("a2a: bad rank/world"); return B200PROBE_EINVAL; }
    if (S & 15) { b200::set_error("a2a: bytes_per_pair must be a multiple of 16"); return B200PROBE_EINVAL; }
    b200::DevProps props;
    int rc = b200::device_props(ordinal, &props);
    if (rc) return rc;
    if (S == 0) return 0;
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    PeerTable t;
    memset(&t, 0, sizeof(t));
    for (int i = 0; i < world; ++i) t.win[i] = (uint8_t*)windows[i];
    const int targets = only_dst >= 0 ? 1 : world;
    if (variant == B200PROBE_VARIANT_TMA) {
        // one CTA per SM in total, split over the destinations; 4 warps x 4 stages x 8 KiB per CTA
        if (ctas_per_peer <= 0) ctas_per_peer = std::max(1, props.sms / targets);
        const uint32_t SB = 8192, NS = 4;
        const size_t smem = (size_t)kRingWarps * NS * SB;
        static bool attr[2] = {false, false};
        // per-device function attribute; cheap enough to set on every launch
        if (sendbuf) B200_CUDA_TRY(cudaFuncSetAttribute(a2a_ring_push_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else B200_CUDA_TRY(cudaFuncSetAttribute(a2a_ring_push_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        (void)attr;
        dim3 g((unsigned)(world * ctas_per_peer));
        if (sendbuf) a2a_ring_push_kernel<true><<<g, kRingWarps * 32, smem, stream>>>(t, (const uint8_t*)sendbuf, rank, world, S, seed, ctas_per_peer, only_dst, SB, NS);
        else a2a_ring_push_kernel<false><<<g, kRingWarps * 32, smem, stream>>>(t, nullptr, rank, world, S, seed, ctas_per_peer, only_dst, SB, NS);
        B200_CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (ctas_per_peer <= 0) ctas_per_peer = default_ctas_per_peer(props.sms, targets);
    dim3 grid((unsigned)(world * ctas_per_peer));
    if (sendbuf) a2a_push_kernel<true><<<grid, 512, 0, stream>>>(t, (const uint8_t*)sendbuf, rank, world, S, seed, ctas_per_peer, only_dst);
    else a2a_push_kernel<false><<<grid, 512, 0, stream>>>(t, nullptr, rank, world, S, seed, ctas_per_peer, only_dst);
    B200_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- NCCL (library leg) -------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
struct Nccl {
    void* lib = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*);
    int (*CommDestroy)(ncclComm_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(int);
};
int load_nccl(Nccl* n) {
    const char* names[] = {getenv("B200PROBE_NCCL_PATH"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        n->lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (n->lib) break;
    }
    if (!n->lib) { b200::set_error("dlopen(libnccl.so.2): %s", dlerror()); return B200PROBE_ENONCCL; }
#define S_(f, name) *(void**)(&n->f) = dlsym(n->lib, name); if (!n->f) { b200::set_error("NCCL symbol %s missing", name); return B200PROBE_ENONCCL; }
    S_(CommInitAll, "ncclCommInitAll") S_(CommDestroy, "ncclCommDestroy") S_(GroupStart, "ncclGroupStart") S_(GroupEnd, "ncclGroupEnd")
    S_(Send, "ncclSend") S_(Recv, "ncclRecv") S_(GetErrorString, "ncclGetErrorString")
#undef S_
    return 0;
}
#define NCCL_TRY(n, expr) do { int r__ = (expr); if (r__ != 0) { b200::set_error("%s -> %s", #expr, (n).GetErrorString(r__)); return B200PROBE_NCCL_BASE + r__; } } while (0)

struct PerDev {
    int ordinal = -1;
    uint8_t *window = nullptr, *sendbuf = nullptr;
    unsigned long long* partials = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

struct A2aCtx {
    std::vector<PerDev> d;
    Nccl nccl;
    std::vector<ncclComm_t> comms;
    ~A2aCtx() {
        for (auto c : comms) if (c && nccl.lib) nccl.CommDestroy(c);
        for (auto& p : d) {
            if (p.ordinal < 0) continue;
            cudaSetDevice(p.ordinal);
            if (p.window) cudaFree(p.window);
            if (p.sendbuf) cudaFree(p.sendbuf);
            if (p.partials) cudaFree(p.partials);
            if (p.e0) cudaEventDestroy(p.e0);
            if (p.e1) cudaEventDestroy(p.e1);
            if (p.stream) cudaStreamDestroy(p.stream);
        }
        if (nccl.lib) dlclose(nccl.lib);
    }
};

double median_of(std::vector<double> v) {
    std::sort(v.begin(), v.end());
    size_t n = v.size();
    return n & 1 ? v[n / 2] : 0.5 * (v[n / 2 - 1] + v[n / 2]);
}

}  // namespace

extern "C" {

uint32_t b200probe_a2a_chunk_seed(uint32_t seed, int src, int dst) { return chunk_seed(seed, src, dst); }

int b200probe_enable_peer_access(const int* ordinals, int g) {
    if (!ordinals || g < 1 || g > kMaxWorld) return B200PROBE_EINVAL;
    for (int i = 0; i < g; ++i) {
        b200::DevProps props;
        int rc = b200::device_props(ordinals[i], &props);
        if (rc) return rc;
        B200_CUDA_TRY(cudaSetDevice(ordinals[i]));
        for (int j = 0; j < g; ++j) {
            if (i == j) continue;
            int can = 0;
            B200_CUDA_TRY(cudaDeviceCanAccessPeer(&can, ordinals[i], ordinals[j]));
            if (!can) { b200::set_error("no peer access %d -> %d", ordinals[i], ordinals[j]); return B200PROBE_ENOPEER; }
            cudaError_t e = cudaDeviceEnablePeerAccess(ordinals[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                b200::set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", ordinals[i], ordinals[j], cudaGetErrorString(e));
                return b200::cuda_rc(e);
            }
            cudaGetLastError();   // clear cudaErrorPeerAccessAlreadyEnabled
        }
    }
    return 0;
}

int b200probe_a2a_window_create(int ordinal, int world, uint64_t S, void** window, unsigned char* handle_out) {
    if (!window || world < 1 || world > kMaxWorld) return B200PROBE_EINVAL;
    b200::DevProps props;
    int rc = b200::device_props(ordinal, &props);
    if (rc) return rc;
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    void* p = nullptr;
    B200_CUDA_TRY(cudaMalloc(&p, std::max<uint64_t>(16, (uint64_t)world * S)));
    B200_CUDA_TRY(cudaMemset(p, 0, std::max<uint64_t>(16, (uint64_t)world * S)));
    if (handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == B200PROBE_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) { cudaFree(p); b200::set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return b200::cuda_rc(e); }
        memcpy(handle_out, &h, sizeof(h));
    }
    *window = p;
    return 0;
}

int b200probe_a2a_window_import(int ordinal, const unsigned char* handle, void** peer_window) {
    if (!handle || !peer_window) return B200PROBE_EINVAL;
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(peer_window, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        b200::set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        return e == cudaErrorPeerAccessUnsupported ? B200PROBE_ENOPEER : b200::cuda_rc(e);
    }
    return 0;
}

int b200probe_a2a_window_release(int ordinal, void* window, int imported) {
    if (!window) return 0;
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    if (imported) B200_CUDA_TRY(cudaIpcCloseMemHandle(window));
    else B200_CUDA_TRY(cudaFree(window));
    return 0;
}

int b200probe_a2a_push(int ordinal, int rank, int world, void* const* windows, uint64_t S, uint32_t seed, int ctas_per_peer, int variant,
                       void* stream) {
    return launch_push(ordinal, rank, world, windows, nullptr, S, seed, ctas_per_peer, -1, variant, (cudaStream_t)stream);
}

int b200probe_a2a_push_buf(int ordinal, int rank, int world, const void* sendbuf, void* const* windows, uint64_t S, int ctas_per_peer,
                           int variant, void* stream) {
    if (!sendbuf) return B200PROBE_EINVAL;
    return launch_push(ordinal, rank, world, windows, sendbuf, S, 0, ctas_per_peer, -1, variant, (cudaStream_t)stream);
}

int b200probe_nvlink_a2a(const int* ordinals, int g, const b200probe_a2a_cfg_t* cfg_in, double* pair_gbs, b200probe_a2a_result_t* out) {
    if (!ordinals || g < 2 || g > kMaxWorld || !out) { b200::set_error("nvlink_a2a: need 2..%d devices", kMaxWorld); return B200PROBE_EINVAL; }
    b200probe_a2a_cfg_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    if (cfg_in) cfg = *cfg_in;
    if (!cfg.bytes_per_pair) cfg.bytes_per_pair = 256ull << 20;
    if (!cfg.warmup && !cfg.reps) { cfg.warmup = 2; cfg.reps = 10; }
    if (cfg.reps < 1) cfg.reps = 1;
    if (!cfg.seed) cfg.seed = 0xB200u;
    const uint64_t S = cfg.bytes_per_pair;
    memset(out, 0, sizeof(*out));
    out->g = g;
    out->verified = -1;
    if (pair_gbs) std::fill(pair_gbs, pair_gbs + g * g, 0.0);

    {
        int rc = b200probe_enable_peer_access(ordinals, g);
        if (rc) return rc;
    }
    A2aCtx ctx;
    ctx.d.resize(g);
    // peer access, windows
    for (int i = 0; i < g; ++i) {
        b200::DevProps props;
        int rc = b200::device_props(ordinals[i], &props);
        if (rc) return rc;
        B200_CUDA_TRY(cudaSetDevice(ordinals[i]));
        PerDev& p = ctx.d[i];
        p.ordinal = ordinals[i];
        B200_CUDA_TRY(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
        B200_CUDA_TRY(cudaEventCreate(&p.e0));
        B200_CUDA_TRY(cudaEventCreate(&p.e1));
        B200_CUDA_TRY(cudaMalloc(&p.window, (size_t)g * S));
        B200_CUDA_TRY(cudaMemset(p.window, 0, (size_t)g * S));
        B200_CUDA_TRY(cudaMalloc(&p.partials, 16));
    }
    void* windows[kMaxWorld] = {nullptr};
    for (int i = 0; i < g; ++i) windows[i] = ctx.d[i].window;

    const bool nccl_mode = cfg.mode == B200PROBE_A2A_NCCL;
    if (nccl_mode) {
        int rc = load_nccl(&ctx.nccl);
        if (rc) return rc;
        ctx.comms.assign(g, nullptr);
        NCCL_TRY(ctx.nccl, ctx.nccl.CommInitAll(ctx.comms.data(), g, ordinals));
        for (int i = 0; i < g; ++i) {   // resident send buffers: chunk for peer p at [p][S]
            PerDev& p = ctx.d[i];
            B200_CUDA_TRY(cudaSetDevice(p.ordinal));
            B200_CUDA_TRY(cudaMalloc(&p.sendbuf, (size_t)g * S));
            for (int j = 0; j < g; ++j) {
                int rc2 = b200probe_hbm_fill(p.ordinal, p.sendbuf + (size_t)j * S, S, chunk_seed(cfg.seed, i, j), nullptr, p.stream);
                if (rc2) return rc2;
            }
            B200_CUDA_TRY(cudaStreamSynchronize(p.stream));
        }
    }

    auto sync_all = [&]() -> int {
        for (int i = 0; i < g; ++i) {
            B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
            B200_CUDA_TRY(cudaStreamSynchronize(ctx.d[i].stream));
        }
        return 0;
    };
    // one all-pairs exchange; per-device elapsed ms into t[]
    auto exchange = [&](bool timed, std::vector<double>* t) -> int {
        if (nccl_mode) {
            for (int i = 0; i < g; ++i) {
                B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
                if (timed) B200_CUDA_TRY(cudaEventRecord(ctx.d[i].e0, ctx.d[i].stream));
            }
            NCCL_TRY(ctx.nccl, ctx.nccl.GroupStart());
            for (int i = 0; i < g; ++i)
                for (int j = 0; j < g; ++j) {
                    if (i == j) continue;
                    NCCL_TRY(ctx.nccl, ctx.nccl.Send(ctx.d[i].sendbuf + (size_t)j * S, S, /*ncclUint8*/ 1, j, ctx.comms[i], ctx.d[i].stream));
                    NCCL_TRY(ctx.nccl, ctx.nccl.Recv(ctx.d[i].window + (size_t)j * S, S, 1, j, ctx.comms[i], ctx.d[i].stream));
                }
            NCCL_TRY(ctx.nccl, ctx.nccl.GroupEnd());
            for (int i = 0; i < g; ++i) {
                B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
                if (timed) B200_CUDA_TRY(cudaEventRecord(ctx.d[i].e1, ctx.d[i].stream));
            }
        } else {
            for (int i = 0; i < g; ++i) {
                PerDev& p = ctx.d[i];
                B200_CUDA_TRY(cudaSetDevice(p.ordinal));
                if (timed) B200_CUDA_TRY(cudaEventRecord(p.e0, p.stream));
                // -2: skip the local slot in the timed exchange (it is HBM traffic, not NVLink)
                int rc = launch_push(p.ordinal, i, g, windows, nullptr, S, cfg.seed, cfg.ctas_per_peer, -2, cfg.variant, p.stream);
                if (rc) return rc;
                if (timed) B200_CUDA_TRY(cudaEventRecord(p.e1, p.stream));
            }
        }
        int rc = sync_all();
        if (rc) return rc;
        if (timed && t) {
            t->assign(g, 0.0);
            for (int i = 0; i < g; ++i) {
                float ms = 0;
                B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
                B200_CUDA_TRY(cudaEventElapsedTime(&ms, ctx.d[i].e0, ctx.d[i].e1));
                (*t)[i] = ms;
            }
        }
        return 0;
    };

    if (cfg.mode == B200PROBE_A2A_PEER_PAIR) {
        double mn = 1e300, mx = 0;
        for (int i = 0; i < g; ++i)
            for (int j = 0; j < g; ++j) {
                if (i == j) continue;
                PerDev& p = ctx.d[i];
                B200_CUDA_TRY(cudaSetDevice(p.ordinal));
                std::vector<double> ts;
                for (int it = -cfg.warmup; it < cfg.reps; ++it) {
                    if (it >= 0) B200_CUDA_TRY(cudaEventRecord(p.e0, p.stream));
                    int rc = launch_push(p.ordinal, i, g, windows, nullptr, S, cfg.seed, cfg.ctas_per_peer, j, cfg.variant, p.stream);
                    if (rc) return rc;
                    if (it >= 0) {
                        float ms;
                        B200_CUDA_TRY(cudaEventRecord(p.e1, p.stream));
                        B200_CUDA_TRY(cudaEventSynchronize(p.e1));
                        B200_CUDA_TRY(cudaEventElapsedTime(&ms, p.e0, p.e1));
                        ts.push_back(ms);
                    }
                }
                B200_CUDA_TRY(cudaStreamSynchronize(p.stream));
                double gbs = (double)S / (median_of(ts) * 1e-3) / 1e9;
                if (pair_gbs) pair_gbs[i * g + j] = gbs;
                mn = std::min(mn, gbs); mx = std::max(mx, gbs);
            }
        out->min_pair_gbs = mn; out->max_pair_gbs = mx;
        // the local slots are not written in pair mode; fill them so verification covers the window
        for (int i = 0; i < g; ++i) {
            int rc = launch_push(ctx.d[i].ordinal, i, g, windows, nullptr, S, cfg.seed, cfg.ctas_per_peer, i, cfg.variant, ctx.d[i].stream);
            if (rc) return rc;
        }
        int rc = sync_all();
        if (rc) return rc;
    } else {
        std::vector<std::vector<double>> per_dev(g);
        std::vector<double> wall;
        for (int it = -cfg.warmup; it < cfg.reps; ++it) {
            std::vector<double> t;
            int rc = exchange(it >= 0, &t);
            if (rc) return rc;
            if (it >= 0) {
                for (int i = 0; i < g; ++i) per_dev[i].push_back(t[i]);
                wall.push_back(*std::max_element(t.begin(), t.end()));
            }
        }
        out->ms_median = median_of(wall);
        out->ms_best = *std::min_element(wall.begin(), wall.end());
        const double payload = (double)(g - 1) * (double)S;
        double mn = 1e300, mx = 0;
        for (int i = 0; i < g; ++i) {
            double ti = median_of(per_dev[i]);
            out->egress_gbs[i] = payload / (ti * 1e-3) / 1e9;
            out->ingress_gbs[i] = payload / (out->ms_median * 1e-3) / 1e9;   // bytes landed over the common window
            for (int j = 0; j < g; ++j) {
                if (i == j) continue;
                double gbs = (double)S / (ti * 1e-3) / 1e9;                  // per-pair share under full concurrency
                if (pair_gbs) pair_gbs[i * g + j] = gbs;
                mn = std::min(mn, gbs); mx = std::max(mx, gbs);
            }
        }
        out->min_pair_gbs = mn; out->max_pair_gbs = mx;
        if (!nccl_mode) {   // local slots (not part of the timed exchange)
            for (int i = 0; i < g; ++i) {
                int rc = launch_push(ctx.d[i].ordinal, i, g, windows, nullptr, S, cfg.seed, cfg.ctas_per_peer, i, cfg.variant, ctx.d[i].stream);
                if (rc) return rc;
            }
            int rc = sync_all();
            if (rc) return rc;
        }
    }

    if (cfg.verify) {
        // expected checksum per (src,dst) chunk, closed form on the host
        int ok = 1;
        for (int dst = 0; dst < g && ok; ++dst) {
            PerDev& p = ctx.d[dst];
            B200_CUDA_TRY(cudaSetDevice(p.ordinal));
            for (int src = 0; src < g; ++src) {
                if (nccl_mode && src == dst) continue;
                B200_CUDA_TRY(cudaMemsetAsync(p.partials, 0, 16, p.stream));
                int rc = b200probe_hbm_read(p.ordinal, p.window + (size_t)src * S, S, (uint64_t*)p.partials, nullptr, p.stream);
                if (rc) return rc;
                unsigned long long h[2];
                B200_CUDA_TRY(cudaMemcpyAsync(h, p.partials, 16, cudaMemcpyDeviceToHost, p.stream));
                B200_CUDA_TRY(cudaStreamSynchronize(p.stream));
                const uint32_t cs = chunk_seed(cfg.seed, src, dst);
                uint64_t es = 0; uint32_t ex = 0;
                for (uint64_t w = 0; w < (S >> 2); ++w) { uint32_t v = b200_pattern_word(w, cs); es += v; ex ^= v; }
                if (h[0] != es || (uint32_t)h[1] != ex) {
                    b200::set_error("a2a: chunk %d->%d landed with checksum %llx/%x, expected %llx/%x", src, dst, h[0], (unsigned)h[1],
                                    (unsigned long long)es, ex);
                    ok = 0;
                    break;
                }
            }
        }
        out->verified = ok;
        if (!ok) return B200PROBE_EMISMATCH;
    }
    return 0;
}

}  // extern "C"
