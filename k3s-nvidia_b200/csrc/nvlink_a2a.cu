// NVLink 5 all-to-all probe (no reference counterpart).
//
// Every rank owns one window in HBM:  [recv: world x S bytes][send: world x S bytes][sync page: 4 KiB]
//   send[p] = the chunk this rank has for rank p   (pattern under chunk_seed(seed, rank, p))
//   recv[p] = where rank p's chunk for this rank lands
// Windows are peer-mapped (cudaDeviceEnablePeerAccess in one process, cudaIpc* across processes)
// and the exchange is done by OUR kernels over those mappings — no NCCL on the data path:
//   PUSH_TMA    pattern generated into shared memory, bulk-STORED (cp.async.bulk, SASS UBLKCP) into the
//               peers' recv slots.  Best when BOTH directions are loaded (the all-to-all): 692 GB/s
//               per direction per GPU on B200 vs 626 for pulls (profiles/a2a_tune_r01_2gpu_variants.txt)
//   PULL_TMA    each rank bulk-LOADS its peers' send chunks over NVLink into shared memory and
//               bulk-stores them into its own recv slots.  Best for ONE direction at a time (the
//               pairwise matrix): 756-781 GB/s vs 699-712 for writes (profiles/p2p_pull_tune_r01.txt)
//   AUTO        (default) PULL_TMA for an isolated pair; for the concurrent exchange PUSH_SYNC when G > 2
//               and S >= 64 MiB (128 MiB without the start gate), else PUSH_TMA
//   PUSH_DIRECT pattern generated in registers, 16-byte stores on the peer pointers
//   PUSH_BUF    local send chunk bulk-loaded, bulk-stored to the peer
//   PUSH_STAGGER PUSH_TMA, but the whole grid works on one peer at a time in the rotation (rank+t) mod G
//   PUSH_SYNC   PUSH_STAGGER with a device-side barrier across ALL ranks before every step (flags in the
//               4 KiB sync page that ends each window, written over NVLink): steps stay aligned, so every
//               GPU sends to one peer and receives from one peer at any time.  701 GB/s per direction at
//               G = 8, S = 256 MiB against 618-660 for the concurrent push (profiles/a2a_sync_relaxed_r01_g8.txt)
//   MIX_TMA     every (src,dst) chunk is moved by BOTH ends at once: the first `split` bytes are pushed
//               by src (PUSH_TMA), the rest pulled by dst (PULL_TMA), so each link direction carries
//               posted writes and read responses side by side (B200PROBE_A2A_MIX_PCT = pushed share)
// CTAs are partitioned per peer so every peer's traffic is in flight at once (NVSwitch gives each
// GPU its full port bandwidth to any mix of peers).  NCCL grouped send/recv is kept as the library
// leg for contrast (mode B200PROBE_A2A_NCCL, dlopen'ed).
//
// Roofline: NVLink bound.  Algorithmic bytes per GPU per direction: (G-1)*S payload bytes.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <pthread.h>

#include <algorithm>
#include <cstddef>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace {

constexpr int kMaxWorld = 16;
constexpr int kRingWarps = 4, kRingMaxStages = 8;

struct PeerTable {
    uint8_t* win[kMaxWorld];   // base of every rank's window (win[rank] is local)
};

B200_HD uint32_t chunk_seed(uint32_t seed, int src, int dst) { return seed ^ b200_mix32((uint32_t)(src * 251 + dst * 7 + 1)); }

struct XArgs {
    PeerTable peers;
    int rank, world;
    uint64_t S;
    uint32_t seed;
    int ctas_per_peer;
    int only_peer;     // >=0 one peer; -1 all slots incl. the local one; -2 all peers, no local slot
    int variant;
    uint32_t SB, NS;   // ring stage bytes / stages
    uint64_t split;    // MIX_TMA: bytes [0,split) of every chunk are pushed by its source, [split,S) pulled by its destination
    int push_ctas;     // MIX_TMA: CTAs (of ctas_per_peer) in the push role
    uint32_t timeout_us;   // PUSH_SYNC: longest wait at a step barrier before the launch gives up synchronising
    int sync_every;        // PUSH_SYNC: a barrier before steps 1, 1+k, 1+2k, ... (k = 1: every step)
    int drain;             // PUSH_SYNC: every warp waits for its bulk stores to COMPLETE before it arrives at a step barrier, and the
                           // last CTA to arrive stamps the step's end: start_ns/done_ns then bracket exactly one pair's bytes on the wire
};

// Sync page at window + 2*world*S (zeroed when the window is created).  flag[q] is written by rank q with a
// remote store over NVLink; cnt and epoch are local.  epoch counts the synchronised launches on this
// window: every rank runs the same sequence of exchanges, so the epochs agree without being communicated,
// and a rank that is behind only makes its peers wait out the barrier timeout (never a hang).
struct SyncPage {
    uint32_t flag[kMaxWorld];
    uint32_t cnt;
    uint32_t epoch;
    // PUSH_SYNC with a barrier before every step.  Step t (1..world-1) moves exactly one pair, rank -> (rank+t) mod G.
    //   start_ns[t]  this device's %globaltimer when CTA 0 left the barrier in front of step t
    //   done_ns[t]   the same clock when the LAST CTA of the grid had seen its bulk stores of step t complete
    //                (cp.async.bulk.wait_group 0).  Written for every step in drain mode, for the last step always.
    // Only a drained step gives a pair rate: without the drain a step "ends" when its stores were ISSUED, and up to
    // SMs x ring bytes (19 MB) of it are still in flight, which made pairs read above the link rate
    // (profiles/a2a_pair_matrix_r01_g4.txt: 1050 GB/s on a 900 GB/s port).
    unsigned long long start_ns[kMaxWorld];
    unsigned long long done_ns[kMaxWorld];
    unsigned int done_cnt;
};
static_assert(sizeof(SyncPage) <= B200PROBE_A2A_SYNC_BYTES, "sync page");

__device__ __forceinline__ bool cta_peer(const XArgs& a, int* peer, int* sub) {
    const int p = blockIdx.x / a.ctas_per_peer;
    if (p >= a.world) return false;
    if (a.only_peer >= 0 && p != a.only_peer) return false;
    if (a.only_peer == -2 && p == a.rank) return false;
    *peer = p;
    *sub = blockIdx.x % a.ctas_per_peer;
    return true;
}

// ---- TMA ring exchange: one ring of shared-memory stages per warp ---------------------------------
__global__ void __launch_bounds__(kRingWarps * 32) a2a_ring_kernel(XArgs a) {
    using namespace b200ptx;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kRingWarps * kRingMaxStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kRingWarps * kRingMaxStages; ++i) mbar_init(smem_u32(&full_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int p, sub;
    if (!cta_peer(a, &p, &sub)) return;
    const uint64_t S = a.S;
    const uint32_t SB = a.SB, NS = a.NS;
    uint8_t* const me = a.peers.win[a.rank];
    uint8_t* const them = a.peers.win[p];
    // role of this CTA and the byte range [lo, hi) of the chunk it works on
    int role = a.variant, nct = a.ctas_per_peer;
    uint64_t lo = 0, hi = S;
    if (a.variant == B200PROBE_A2A_MIX_TMA) {
        if (sub < a.push_ctas) { role = B200PROBE_A2A_PUSH_TMA; nct = a.push_ctas; hi = a.split; }
        else { role = B200PROBE_A2A_PULL_TMA; sub -= a.push_ctas; nct = a.ctas_per_peer - a.push_ctas; lo = a.split; }
        if (p == a.rank) {                              // local slot (only_peer == -1): one role writes all of it
            const bool mine = a.push_ctas > 0 ? role == B200PROBE_A2A_PUSH_TMA : true;
            lo = 0; hi = mine ? S : 0;
        }
    }
    const uint8_t* in;
    uint8_t* out;
    if (role == B200PROBE_A2A_PULL_TMA) {               // their send[rank] -> my recv[p]
        in = them + ((uint64_t)a.world + a.rank) * S;
        out = me + (uint64_t)p * S;
    } else {                                            // (my send[p] | generated) -> their recv[rank]
        in = me + ((uint64_t)a.world + p) * S;
        out = them + (uint64_t)a.rank * S;
    }
    const uint64_t nchunks = hi > lo ? (hi - lo + SB - 1) / SB : 0;
    const uint64_t worker = (uint64_t)sub * kRingWarps + warp, nworkers = (uint64_t)nct * kRingWarps;
    const uint64_t n_my = worker < nchunks ? (nchunks - worker + nworkers - 1) / nworkers : 0;
    const uint32_t ring = smem_u32(smem) + warp * NS * SB;
    uint8_t* ring_ptr = smem + (size_t)warp * NS * SB;
    const uint32_t bar0 = smem_u32(&full_bar[warp * kRingMaxStages]);
    const uint64_t pol = policy_evict_first();
    auto off_of = [&](uint64_t k) { return lo + (worker + k * nworkers) * (uint64_t)SB; };
    auto len_of = [&](uint64_t k) { return (uint32_t)min((uint64_t)SB, hi - off_of(k)); };
    uint32_t cs = 0, cph = 0;
    if (role != B200PROBE_A2A_PUSH_TMA) {
        if (lane == 0 && n_my > 0) {       // bulk load -> bulk store; the data never touches registers
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
                mbar_wait(bar0 + cs * 8, cph);
                bulk_s2g(out + off_of(k), ring + cs * SB, len_of(k), pol);
                bulk_commit();
                if (issued < n_my) { bulk_wait_read<1>(); load_next(); }
                if (++cs == NS) { cs = 0; cph ^= 1; }
            }
            bulk_wait_all();
        }
    } else {
        const uint32_t sd = chunk_seed(a.seed, a.rank, p);
        for (uint64_t k = 0; k < n_my; ++k) {
            if (k >= NS) {
                if (lane == 0) bulk_wait_read_dyn((int)NS - 1);
                __syncwarp();
            }
            uint4* st = reinterpret_cast<uint4*>(ring_ptr + (size_t)cs * SB);
            const uint32_t nvec = len_of(k) >> 4;
            const uint64_t w0 = off_of(k) >> 2;
#pragma unroll 4
            for (uint32_t i = lane; i < nvec; i += 32) {
                const uint64_t w = w0 + (uint64_t)i * 4;
                st[i] = make_uint4(b200_pattern_word(w, sd), b200_pattern_word(w + 1, sd), b200_pattern_word(w + 2, sd), b200_pattern_word(w + 3, sd));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_s2g(out + off_of(k), ring + cs * SB, len_of(k), pol); bulk_commit(); }
            if (++cs == NS) cs = 0;
        }
        if (lane == 0) bulk_wait_all();
    }
}

// ---- staggered push: the whole grid visits the peers ONE AT A TIME in the rotation (rank+t) mod world,
// so at any moment every GPU receives from a single source (no output-port sharing inside NVSwitch).
// Same data path as PUSH_TMA (pattern -> shared memory -> bulk store); the ring of stages runs on
// across peers without draining.
__global__ void __launch_bounds__(kRingWarps * 32) a2a_stagger_kernel(XArgs a) {
    using namespace b200ptx;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t S = a.S;
    const uint32_t SB = a.SB, NS = a.NS;
    const uint64_t nchunks = (S + SB - 1) / SB;
    const uint64_t worker = (uint64_t)blockIdx.x * kRingWarps + warp, nworkers = (uint64_t)gridDim.x * kRingWarps;
    const uint64_t n_my = worker < nchunks ? (nchunks - worker + nworkers - 1) / nworkers : 0;
    const uint32_t ring = smem_u32(smem) + warp * NS * SB;
    uint8_t* ring_ptr = smem + (size_t)warp * NS * SB;
    const uint64_t pol = policy_evict_first();
    uint32_t cs = 0;
    uint64_t stored = 0;
    const bool sync = a.variant == B200PROBE_A2A_PUSH_SYNC && a.only_peer < 0;
    __shared__ int s_desync;
    if (threadIdx.x == 0) s_desync = 0;
    SyncPage* const mine = reinterpret_cast<SyncPage*>(a.peers.win[a.rank] + 2ull * a.world * S);
    uint32_t epoch = 0;
    if (sync) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(epoch) : "l"(&mine->epoch) : "memory");   // bumped only after every CTA's last arrival
    for (int t = (a.only_peer == -1 ? 0 : 1); t < a.world; ++t) {
        const int p = (a.rank + t) % a.world;
        if (a.only_peer >= 0 && p != a.only_peer) continue;
        if (sync && t >= 1 && (t - 1) % a.sync_every == 0) {
            // ---- barrier t-1 over every CTA of every rank: arrive locally, last arriver raises this rank's
            // flag on all ranks, then everyone polls its LOCAL sync page.  A wait longer than timeout_us gives
            // up synchronising for the rest of the launch (a peer that never launched must not hang the GPU).
            const uint32_t b = (uint32_t)((t - 1) / a.sync_every), nb = (uint32_t)((a.world - 2) / a.sync_every + 1);
            const uint32_t target = epoch * 16u + b + 1u;
            if (a.drain && t >= 2) {            // step t-1 is over for this warp only when its stores have landed
                if (lane == 0) bulk_wait_all();
                __syncwarp();
            }
            __syncthreads();                    // every warp of this CTA has issued (drain: completed) its stores of the step before
            if (threadIdx.x == 0) {
                const uint32_t n = atomicAdd(&mine->cnt, 1u);
                if (n == gridDim.x * (b + 1) - 1) {
                    if (a.drain && t >= 2) {    // last CTA of this rank to finish step t-1
                        unsigned long long now_ns;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_ns));
                        mine->done_ns[t - 1] = now_ns;
                    }
                    if (b + 1 == nb) { atomicExch(&mine->cnt, 0u); atomicExch(&mine->epoch, epoch + 1u); }   // last barrier of the launch
                    for (int q = 0; q < a.world; ++q) {
                        SyncPage* theirs = reinterpret_cast<SyncPage*>(a.peers.win[q] + 2ull * a.world * S);
                        // a pacing signal, not a publication: relaxed.  (st.release.sys fences the SM's outstanding
                        // stores first — with megabytes of bulk stores in flight that drained the pipeline at every step.)
                        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(&theirs->flag[a.rank]), "r"(target) : "memory");
                    }
                }
            }
            if (warp == 0 && !s_desync) {
                uint64_t t0, now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                bool ok = false;
                while (!ok) {
                    uint32_t v = target;
                    if (lane < a.world) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(&mine->flag[lane]) : "memory");
                    ok = __all_sync(0xffffffffu, (int32_t)(v - target) >= 0);
                    if (!ok) {
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (__any_sync(0xffffffffu, now - t0 > (uint64_t)a.timeout_us * 1000ull)) {      // warp-uniform decision
                            if (lane == 0) s_desync = 1;
                            break;
                        }
                        __nanosleep(200);
                    }
                }
                if (blockIdx.x == 0 && lane == 0 && a.sync_every == 1) {
                    unsigned long long now_ns;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_ns));
                    mine->start_ns[t] = ok ? now_ns : 0ull;      // 0: this launch lost the barrier, its stamps mean nothing
                }
            }
            __syncthreads();
        }
        uint8_t* const out = a.peers.win[p] + (uint64_t)a.rank * S;
        const uint32_t sd = chunk_seed(a.seed, a.rank, p);
        for (uint64_t k = 0; k < n_my; ++k, ++stored) {
            const uint64_t off = (worker + k * nworkers) * (uint64_t)SB;
            const uint32_t len = (uint32_t)min((uint64_t)SB, S - off);
            if (stored >= NS) {
                if (lane == 0) bulk_wait_read_dyn((int)NS - 1);
                __syncwarp();
            }
            uint4* st = reinterpret_cast<uint4*>(ring_ptr + (size_t)cs * SB);
            const uint32_t nvec = len >> 4;
            const uint64_t w0 = off >> 2;
#pragma unroll 4
            for (uint32_t i = lane; i < nvec; i += 32) {
                const uint64_t w = w0 + (uint64_t)i * 4;
                st[i] = make_uint4(b200_pattern_word(w, sd), b200_pattern_word(w + 1, sd), b200_pattern_word(w + 2, sd), b200_pattern_word(w + 3, sd));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_s2g(out + off, ring + cs * SB, len, pol); bulk_commit(); }
            if (++cs == NS) cs = 0;
        }
    }
    if (lane == 0) bulk_wait_all();
    if (sync && a.sync_every == 1) {
        __syncthreads();                                   // every warp of this CTA has seen its stores complete
        if (threadIdx.x == 0) {
            if (atomicAdd(&mine->done_cnt, 1u) == gridDim.x - 1) {          // last CTA of the grid: the last step (and the exchange) is over
                unsigned long long now_ns;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_ns));
                mine->done_ns[a.world - 1] = now_ns;
                atomicExch(&mine->done_cnt, 0u);
            }
        }
    }
}

// ---- start gate (single-process probe): every device's exchange kernel is queued behind one of these and
// all of them are released by ONE host store, so the exchanges start together instead of one launch overhead
// apart (8 sequential launches are ~100 us of skew).  Spins on a word in mapped pinned host memory.
__global__ void a2a_gate_kernel(const volatile uint32_t* gate, uint32_t want) {
    while ((int32_t)(*gate - want) < 0) __nanosleep(100);       // released once the host word has reached this exchange's number
}

// ---- direct push: pattern in registers, 16-byte stores on the peer pointer --------------------------
__global__ void __launch_bounds__(512) a2a_direct_kernel(XArgs a) {
    int p, sub;
    if (!cta_peer(a, &p, &sub)) return;
    uint4* __restrict__ out = reinterpret_cast<uint4*>(a.peers.win[p] + (uint64_t)a.rank * a.S);
    const uint32_t sd = chunk_seed(a.seed, a.rank, p);
    const uint64_t nvec = a.S >> 4;
    const uint64_t stride = (uint64_t)a.ctas_per_peer * blockDim.x;
    uint64_t i = (uint64_t)sub * blockDim.x + threadIdx.x;
    constexpr int U = 8;
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t w = (i + j * stride) * 4;
            v[j] = make_uint4(b200_pattern_word(w, sd), b200_pattern_word(w + 1, sd), b200_pattern_word(w + 2, sd), b200_pattern_word(w + 3, sd));
        }
#pragma unroll
        for (int j = 0; j < U; ++j) out[i + j * stride] = v[j];
    }
    for (; i < nvec; i += stride) {
        const uint64_t w = i * 4;
        out[i] = make_uint4(b200_pattern_word(w, sd), b200_pattern_word(w + 1, sd), b200_pattern_word(w + 2, sd), b200_pattern_word(w + 3, sd));
    }
}

// AUTO for the concurrent exchange.  With the relaxed-flag barrier a step barrier costs a few microseconds, so
// keeping the steps aligned wins as soon as a step is long against the start skew between ranks: at G = 8
// from S = 64 MiB behind the start gate of the single-process probe (681 vs 655 GB/s; 701 vs 618 at 256 MiB,
// profiles/a2a_sync_relaxed_r01_g8.txt), one size later for free-running launches.  Below that the
// concurrent push is as good or better (8 MiB: 604 vs 579).  Two ranks: one peer, nothing to stagger.
int auto_exchange_variant(int world, uint64_t S, bool gated = false) {
    return (world > 2 && S >= ((gated ? 64ull : 128ull) << 20)) ? B200PROBE_A2A_PUSH_SYNC : B200PROBE_A2A_PUSH_TMA;
}

int launch_exchange(int ordinal, int rank, int world, void* const* windows, uint64_t S, uint32_t seed, int variant, int ctas_per_peer,
                    int only_peer, cudaStream_t stream, bool drain = false) {
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || !windows) { b200::set_error("a2a: bad rank/world"); return B200PROBE_EINVAL; }
    if (S & 15) { b200::set_error("a2a: bytes_per_pair must be a multiple of 16"); return B200PROBE_EINVAL; }
    if (variant < 0 || variant > B200PROBE_A2A_PUSH_SYNC) { b200::set_error("a2a: unknown variant %d", variant); return B200PROBE_EINVAL; }
    if (variant == B200PROBE_A2A_AUTO) variant = only_peer >= 0 ? B200PROBE_A2A_PULL_TMA : auto_exchange_variant(world, S);
    if (variant == B200PROBE_A2A_MIX_TMA && only_peer >= 0) variant = B200PROBE_A2A_PULL_TMA;   // one direction at a time: nothing to mix
    if (only_peer >= world) { b200::set_error("a2a: peer %d out of range", only_peer); return B200PROBE_ERANGE; }
    b200::DevProps props;
    int rc = b200::device_props(ordinal, &props);
    if (rc) return rc;
    if (S == 0) return 0;
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    XArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < world; ++i) a.peers.win[i] = (uint8_t*)windows[i];
    a.rank = rank; a.world = world; a.S = S; a.seed = seed; a.only_peer = only_peer; a.variant = variant;
    const int targets = only_peer >= 0 ? 1 : std::max(1, world - (only_peer == -2 ? 1 : 0));
    if (variant == B200PROBE_A2A_PUSH_DIRECT) {
        a.ctas_per_peer = ctas_per_peer > 0 ? ctas_per_peer : std::max(1, (props.sms * 4 + targets - 1) / targets);
        a2a_direct_kernel<<<world * a.ctas_per_peer, 512, 0, stream>>>(a);
    } else {
        // about one CTA per SM in total, split over the peers; 4 warps x 4 stages x 8 KiB = 128 KiB per CTA
        a.ctas_per_peer = ctas_per_peer > 0 ? ctas_per_peer : std::max(1, props.sms / targets);
        a.SB = 8192; a.NS = 4;
        if (const char* e = getenv("B200PROBE_A2A_STAGE_BYTES")) { int v = atoi(e); if (v >= 512 && v <= 49152 && !(v & 511)) a.SB = (uint32_t)v; }
        if (const char* e = getenv("B200PROBE_A2A_STAGES")) { int v = atoi(e); if (v >= 2 && v <= kRingMaxStages) a.NS = (uint32_t)v; }
        if ((size_t)kRingWarps * a.NS * a.SB > 200u * 1024u) { b200::set_error("a2a: ring of %u x %u B stages does not fit shared memory", a.NS, a.SB); return B200PROBE_EINVAL; }
        if (variant == B200PROBE_A2A_MIX_TMA) {
            int pct = 50;
            if (const char* e = getenv("B200PROBE_A2A_MIX_PCT")) pct = std::min(100, std::max(0, atoi(e)));
            const uint64_t stages_total = (S + a.SB - 1) / a.SB;
            a.split = std::min<uint64_t>(S, (stages_total * pct / 100) * a.SB);
            if (a.ctas_per_peer < 2) a.ctas_per_peer = 2;
            a.push_ctas = a.split == 0 ? 0 : a.split == S ? a.ctas_per_peer
                                                          : std::min(a.ctas_per_peer - 1, std::max(1, (int)((int64_t)a.ctas_per_peer * pct / 100)));
        }
        const size_t smem = (size_t)kRingWarps * a.NS * a.SB;
        if (variant == B200PROBE_A2A_PUSH_STAGGER || variant == B200PROBE_A2A_PUSH_SYNC) {
            int grid = ctas_per_peer > 0 ? ctas_per_peer : props.sms;            // here: CTAs in total, all on one peer at a time
            if (variant == B200PROBE_A2A_PUSH_SYNC) {
                grid = std::min(grid, props.sms);                                // the barrier needs every CTA resident (1 CTA/SM at this ring size)
                a.timeout_us = 200000;
                a.sync_every = 1;
                a.drain = drain ? 1 : 0;
                if (const char* e = getenv("B200PROBE_A2A_SYNC_EVERY")) { int v = atoi(e); if (v >= 1 && v <= kMaxWorld) a.sync_every = v; }
                if (const char* e = getenv("B200PROBE_A2A_SYNC_TIMEOUT_US")) { int v = atoi(e); if (v > 0) a.timeout_us = (uint32_t)v; }
            }
            B200_CUDA_TRY(cudaFuncSetAttribute(a2a_stagger_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            a2a_stagger_kernel<<<grid, kRingWarps * 32, smem, stream>>>(a);
        } else {
            B200_CUDA_TRY(cudaFuncSetAttribute(a2a_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            a2a_ring_kernel<<<world * a.ctas_per_peer, kRingWarps * 32, smem, stream>>>(a);
        }
    }
    B200_CUDA_TRY(cudaGetLastError());
    return 0;
}

int fill_send_half(int ordinal, void* window, int rank, int world, uint64_t S, uint32_t seed, cudaStream_t stream) {
    for (int p = 0; p < world; ++p) {
        int rc = b200probe_hbm_fill(ordinal, (uint8_t*)window + ((uint64_t)world + p) * S, S, chunk_seed(seed, rank, p), nullptr, stream);
        if (rc) return rc;
    }
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
    uint8_t* window = nullptr;
    unsigned long long* partials = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

// Resident per-process exchange context: windows (send halves filled), streams, events, the start gate and the NCCL
// communicators of the library leg stay alive between calls with the same (devices, S, seed) — a probe round calls the
// entry several times (exchange, pair matrix, library contrast) and cudaMalloc of 4 GiB per device plus ncclCommInitAll
// cost far more than the exchange.  Freed by b200probe_a2a_release(), or replaced when the key changes.
struct A2aCtx {
    int g = 0;
    int ordinals[kMaxWorld] = {0};
    uint64_t S = 0;
    uint32_t seed = 0;
    std::vector<PerDev> d;
    uint32_t* gate_host = nullptr;      // mapped pinned word the gate kernels spin on
    uint32_t gate_epoch = 0;
    Nccl nccl;
    std::vector<ncclComm_t> comms;
    bool matches(const int* o, int n, uint64_t s, uint32_t sd) const {
        if (n != g || s != S || sd != seed) return false;
        for (int i = 0; i < n; ++i) if (o[i] != ordinals[i]) return false;
        return true;
    }
    ~A2aCtx() {
        for (auto c : comms) if (c && nccl.lib) nccl.CommDestroy(c);
        for (auto& p : d) {
            if (p.ordinal < 0) continue;
            cudaSetDevice(p.ordinal);
            if (p.window) cudaFree(p.window);
            if (p.partials) cudaFree(p.partials);
            if (p.e0) cudaEventDestroy(p.e0);
            if (p.e1) cudaEventDestroy(p.e1);
            if (p.stream) cudaStreamDestroy(p.stream);
        }
        if (nccl.lib) dlclose(nccl.lib);
        if (gate_host) cudaFreeHost(gate_host);
    }
};
pthread_mutex_t g_ctx_mu = PTHREAD_MUTEX_INITIALIZER;     // also serialises the cross-device calls (header: "serialise internally")
A2aCtx* g_ctx = nullptr;


int ctx_build(A2aCtx* c, const int* ordinals, int g, uint64_t S, uint32_t seed) {
    c->g = g; c->S = S; c->seed = seed;
    for (int i = 0; i < g; ++i) c->ordinals[i] = ordinals[i];
    c->d.resize(g);
    for (int i = 0; i < g; ++i) {
        B200_CUDA_TRY(cudaSetDevice(ordinals[i]));
        PerDev& p = c->d[i];
        p.ordinal = ordinals[i];
        B200_CUDA_TRY(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
        B200_CUDA_TRY(cudaEventCreate(&p.e0));
        B200_CUDA_TRY(cudaEventCreate(&p.e1));
        B200_ALLOC_TRY(cudaMalloc(&p.window, 2ull * g * S + B200PROBE_A2A_SYNC_BYTES));
        B200_ALLOC_TRY(cudaMalloc(&p.partials, 32));
        int rc = fill_send_half(p.ordinal, p.window, i, g, S, seed, p.stream);
        if (rc) return rc;
    }
    B200_CUDA_TRY(cudaHostAlloc((void**)&c->gate_host, 64, cudaHostAllocPortable | cudaHostAllocMapped));
    *c->gate_host = 0;
    for (int i = 0; i < g; ++i) {
        B200_CUDA_TRY(cudaSetDevice(ordinals[i]));
        B200_CUDA_TRY(cudaStreamSynchronize(c->d[i].stream));
    }
    return 0;
}

// the context for this key, built or reused; g_ctx_mu held by the caller
int ctx_acquire(const int* ordinals, int g, uint64_t S, uint32_t seed, A2aCtx** out) {
    if (g_ctx && !g_ctx->matches(ordinals, g, S, seed)) { delete g_ctx; g_ctx = nullptr; }
    if (!g_ctx) {
        A2aCtx* c = new A2aCtx();
        int rc = ctx_build(c, ordinals, g, S, seed);
        if (rc) { delete c; return rc; }
        g_ctx = c;
    }
    *out = g_ctx;
    return 0;
}

struct MutexLock {
    pthread_mutex_t* m;
    explicit MutexLock(pthread_mutex_t* x) : m(x) { pthread_mutex_lock(m); }
    ~MutexLock() { pthread_mutex_unlock(m); }
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
    static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    static bool enabled[B200PROBE_MAX_DEVICES][B200PROBE_MAX_DEVICES];      // per process: asking twice is an API error, so remember
    struct Lock { pthread_mutex_t* m; explicit Lock(pthread_mutex_t* x) : m(x) { pthread_mutex_lock(m); } ~Lock() { pthread_mutex_unlock(m); } } lock(&mu);
    for (int i = 0; i < g; ++i) {
        b200::DevProps props;
        int rc = b200::device_props(ordinals[i], &props);
        if (rc) return rc;
        if (ordinals[i] < 0 || ordinals[i] >= B200PROBE_MAX_DEVICES) return B200PROBE_ERANGE;
        B200_CUDA_TRY(cudaSetDevice(ordinals[i]));
        for (int j = 0; j < g; ++j) {
            if (i == j || ordinals[j] < 0 || ordinals[j] >= B200PROBE_MAX_DEVICES || enabled[ordinals[i]][ordinals[j]]) continue;
            int can = 0;
            B200_CUDA_TRY(cudaDeviceCanAccessPeer(&can, ordinals[i], ordinals[j]));
            if (!can) { b200::set_error("no peer access %d -> %d", ordinals[i], ordinals[j]); return B200PROBE_ENOPEER; }
            cudaError_t e = cudaDeviceEnablePeerAccess(ordinals[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {       // "already": another library in this process (torch) got there first
                b200::set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", ordinals[i], ordinals[j], cudaGetErrorString(e));
                return b200::cuda_rc(e);
            }
            if (e != cudaSuccess) cudaGetLastError();
            enabled[ordinals[i]][ordinals[j]] = true;
        }
    }
    return 0;
}

int b200probe_a2a_window_create(int ordinal, int world, uint64_t S, void** window, unsigned char* handle_out) {
    if (!window || world < 1 || world > kMaxWorld || (S & 15)) return B200PROBE_EINVAL;
    b200::DevProps props;
    int rc = b200::device_props(ordinal, &props);
    if (rc) return rc;
    B200_CUDA_TRY(cudaSetDevice(ordinal));
    void* p = nullptr;
    const uint64_t bytes = 2ull * world * S + B200PROBE_A2A_SYNC_BYTES;
    B200_CUDA_TRY(cudaMalloc(&p, bytes));
    B200_CUDA_TRY(cudaMemset(p, 0, bytes));
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

int b200probe_a2a_window_fill(int ordinal, void* window, int rank, int world, uint64_t S, uint32_t seed, void* stream) {
    if (!window || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || (S & 15)) return B200PROBE_EINVAL;
    return fill_send_half(ordinal, window, rank, world, S, seed, (cudaStream_t)stream);
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

int b200probe_a2a_exchange(int ordinal, int rank, int world, void* const* windows, uint64_t S, uint32_t seed, int variant, int ctas_per_peer,
                           int only_peer, void* stream) {
    return launch_exchange(ordinal, rank, world, windows, S, seed, variant, ctas_per_peer, only_peer, (cudaStream_t)stream);
}

int b200probe_a2a_release(void) {
    MutexLock lock(&g_ctx_mu);
    delete g_ctx;
    g_ctx = nullptr;
    return 0;
}

static int nvlink_a2a_locked(const int* ordinals, int g, const b200probe_a2a_cfg_t* cfg_in, double* pair_gbs, b200probe_a2a_result_t* out) {
    if (!ordinals || g < 2 || g > kMaxWorld || !out) { b200::set_error("nvlink_a2a: need 2..%d devices", kMaxWorld); return B200PROBE_EINVAL; }
    b200probe_a2a_cfg_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    if (cfg_in) cfg = *cfg_in;
    if (!cfg.bytes_per_pair) cfg.bytes_per_pair = 256ull << 20;
    if (!cfg.warmup && !cfg.reps) { cfg.warmup = 2; cfg.reps = 10; }
    if (cfg.reps < 1) cfg.reps = 1;
    if (!cfg.seed) cfg.seed = 0xB200u;
    const uint64_t S = cfg.bytes_per_pair;
    if (S & 15) { b200::set_error("a2a: bytes_per_pair must be a multiple of 16"); return B200PROBE_EINVAL; }
    if (cfg.mode < B200PROBE_A2A_PEER_ALL || cfg.mode > B200PROBE_A2A_CE) { b200::set_error("a2a: unknown mode %d", cfg.mode); return B200PROBE_EINVAL; }
    memset(out, 0, sizeof(*out));
    out->g = g;
    out->verified = -1;
    if (pair_gbs) std::fill(pair_gbs, pair_gbs + g * g, 0.0);
    const bool nccl_mode = cfg.mode == B200PROBE_A2A_NCCL, ce_mode = cfg.mode == B200PROBE_A2A_CE;
    const bool library = nccl_mode || ce_mode;
    const bool gated = !nccl_mode && cfg.mode != B200PROBE_A2A_PEER_PAIR && !getenv("B200PROBE_A2A_NO_GATE");
    if (cfg.variant == B200PROBE_A2A_AUTO) cfg.variant = cfg.mode == B200PROBE_A2A_PEER_PAIR ? B200PROBE_A2A_PULL_TMA : auto_exchange_variant(g, cfg.bytes_per_pair, gated);
    if (cfg.variant == B200PROBE_A2A_MIX_TMA && cfg.mode == B200PROBE_A2A_PEER_PAIR) cfg.variant = B200PROBE_A2A_PULL_TMA;
    const bool pull = !library && cfg.variant == B200PROBE_A2A_PULL_TMA;
    const bool mix = !library && cfg.variant == B200PROBE_A2A_MIX_TMA;

    int rc = b200probe_enable_peer_access(ordinals, g);
    if (rc) return rc;
    A2aCtx* ctxp = nullptr;
    rc = ctx_acquire(ordinals, g, S, cfg.seed, &ctxp);
    if (rc) return rc;
    A2aCtx& ctx = *ctxp;
    // every call starts from empty recv slots (what verification finds there landed during THIS call) and a clean sync page
    for (int i = 0; i < g; ++i) {
        B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
        B200_CUDA_TRY(cudaMemsetAsync(ctx.d[i].window, 0, (size_t)g * S, ctx.d[i].stream));
        B200_CUDA_TRY(cudaMemsetAsync(ctx.d[i].window + 2ull * g * S, 0, B200PROBE_A2A_SYNC_BYTES, ctx.d[i].stream));
    }
    void* windows[kMaxWorld] = {nullptr};
    for (int i = 0; i < g; ++i) windows[i] = ctx.d[i].window;

    if (nccl_mode && ctx.comms.empty()) {
        rc = load_nccl(&ctx.nccl);
        if (rc) return rc;
        ctx.comms.assign(g, nullptr);
        int r = ctx.nccl.CommInitAll(ctx.comms.data(), g, ordinals);
        if (r != 0) { ctx.comms.clear(); b200::set_error("ncclCommInitAll -> %s", ctx.nccl.GetErrorString(r)); return B200PROBE_NCCL_BASE + r; }
    }

    auto sync_all = [&]() -> int {
        for (int i = 0; i < g; ++i) {
            B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
            B200_CUDA_TRY(cudaStreamSynchronize(ctx.d[i].stream));
        }
        return 0;
    };
    rc = sync_all();
    if (rc) return rc;
    // one all-pairs exchange; per-device elapsed ms into t[]
    auto exchange = [&](bool timed, std::vector<double>* t, bool drain) -> int {
        const uint32_t release = ++ctx.gate_epoch;
        struct Release {                                   // whatever path leaves this function, no gate kernel is left spinning
            uint32_t* word; uint32_t v;
            ~Release() { if (word) __atomic_store_n(word, v, __ATOMIC_RELEASE); }
        } release_guard{gated ? ctx.gate_host : nullptr, release};
        for (int i = 0; i < g; ++i) {
            B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
            if (gated) {
                uint32_t* dev_gate = nullptr;
                B200_CUDA_TRY(cudaHostGetDevicePointer((void**)&dev_gate, ctx.gate_host, 0));
                a2a_gate_kernel<<<1, 1, 0, ctx.d[i].stream>>>(dev_gate, release);
                B200_CUDA_TRY(cudaGetLastError());
            }
            if (timed) B200_CUDA_TRY(cudaEventRecord(ctx.d[i].e0, ctx.d[i].stream));
            if (ce_mode) {
                // library leg 2: the copy engines push every chunk (cudaMemcpyPeerAsync), peers in the rotation (i+t) mod G
                for (int t = 1; t < g; ++t) {
                    const int j = (i + t) % g;
                    B200_CUDA_TRY(cudaMemcpyPeerAsync(ctx.d[j].window + (size_t)i * S, ctx.d[j].ordinal, ctx.d[i].window + ((size_t)g + j) * S,
                                                      ctx.d[i].ordinal, S, ctx.d[i].stream));
                }
                if (timed) B200_CUDA_TRY(cudaEventRecord(ctx.d[i].e1, ctx.d[i].stream));
            } else if (!nccl_mode) {
                // -2: the local slot is HBM traffic, not NVLink: kept out of the timed exchange
                int r2 = launch_exchange(ctx.d[i].ordinal, i, g, windows, S, cfg.seed, cfg.variant, cfg.ctas_per_peer, -2, ctx.d[i].stream, drain);
                if (r2) return r2;
                if (timed) B200_CUDA_TRY(cudaEventRecord(ctx.d[i].e1, ctx.d[i].stream));
            }
        }
        if (gated) __atomic_store_n(ctx.gate_host, release, __ATOMIC_RELEASE);      // every device's exchange starts now
        if (nccl_mode) {
            NCCL_TRY(ctx.nccl, ctx.nccl.GroupStart());
            for (int i = 0; i < g; ++i)
                for (int j = 0; j < g; ++j) {
                    if (i == j) continue;
                    NCCL_TRY(ctx.nccl, ctx.nccl.Send(ctx.d[i].window + ((size_t)g + j) * S, S, /*ncclUint8*/ 1, j, ctx.comms[i], ctx.d[i].stream));
                    NCCL_TRY(ctx.nccl, ctx.nccl.Recv(ctx.d[i].window + (size_t)j * S, S, 1, j, ctx.comms[i], ctx.d[i].stream));
                }
            NCCL_TRY(ctx.nccl, ctx.nccl.GroupEnd());
            for (int i = 0; i < g; ++i) {
                B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
                if (timed) B200_CUDA_TRY(cudaEventRecord(ctx.d[i].e1, ctx.d[i].stream));
            }
        }
        int r3 = sync_all();
        if (r3) return r3;
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
    auto local_slots = [&]() -> int {       // own chunk into own recv slot (not NVLink traffic; completes the window)
        for (int i = 0; i < g; ++i) {
            B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
            B200_CUDA_TRY(cudaMemcpyAsync(ctx.d[i].window + (size_t)i * S, ctx.d[i].window + ((size_t)g + i) * S, S, cudaMemcpyDeviceToDevice, ctx.d[i].stream));
        }
        return sync_all();
    };

    if (cfg.mode == B200PROBE_A2A_PEER_PAIR) {
        double mn = 1e300, mx = 0;
        for (int src = 0; src < g; ++src)
            for (int dst = 0; dst < g; ++dst) {
                if (src == dst) continue;
                const int runner = pull ? dst : src, peer = pull ? src : dst;    // the rank whose kernel moves src -> dst
                PerDev& p = ctx.d[runner];
                B200_CUDA_TRY(cudaSetDevice(p.ordinal));
                std::vector<double> ts;
                for (int it = -cfg.warmup; it < cfg.reps; ++it) {
                    if (it >= 0) B200_CUDA_TRY(cudaEventRecord(p.e0, p.stream));
                    rc = launch_exchange(p.ordinal, runner, g, windows, S, cfg.seed, cfg.variant, cfg.ctas_per_peer, peer, p.stream);
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
                const double gbs = (double)S / (median_of(ts) * 1e-3) / 1e9;
                if (pair_gbs) pair_gbs[src * g + dst] = gbs;
                mn = std::min(mn, gbs); mx = std::max(mx, gbs);
            }
        out->min_pair_gbs = mn; out->max_pair_gbs = mx;
        out->pair_source = B200PROBE_PAIR_ISOLATED;
    } else {
        std::vector<std::vector<double>> per_dev(g);
        std::vector<double> wall;
        for (int it = -cfg.warmup; it < cfg.reps; ++it) {
            std::vector<double> t;
            rc = exchange(it >= 0, &t, false);
            if (rc) return rc;
            if (it >= 0) {
                for (int i = 0; i < g; ++i) per_dev[i].push_back(t[i]);
                wall.push_back(*std::max_element(t.begin(), t.end()));
            }
        }
        out->ms_median = median_of(wall);
        out->ms_best = *std::min_element(wall.begin(), wall.end());
        const double payload = (double)(g - 1) * (double)S;
        // Pair matrix of the stepped exchange (PUSH_SYNC, a barrier before every step): a step moves one pair per rank alone
        // (rank -> (rank+t) mod G).  The timed exchanges above run the steps back to back WITHOUT draining (that is the
        // headline rate); the matrix comes from extra exchanges in drain mode, where a step starts when every rank's previous
        // step has landed and ends when this rank's last store has completed: S bytes / (done - start) is that pair's own
        // rate with nothing else of this rank on the wire, and it cannot exceed the port rate.
        std::vector<std::vector<std::vector<double>>> stepped;      // [i][t] -> GB/s samples
        if (!library && cfg.variant == B200PROBE_A2A_PUSH_SYNC && g > 2 && !getenv("B200PROBE_A2A_SYNC_EVERY")) {
            stepped.assign(g, std::vector<std::vector<double>>(g));
            const int kMatrixReps = 3;
            for (int rep = 0; rep < kMatrixReps && !stepped.empty(); ++rep) {
                rc = exchange(false, nullptr, true);
                if (rc) return rc;
                for (int i = 0; i < g && !stepped.empty(); ++i) {
                    B200_CUDA_TRY(cudaSetDevice(ctx.d[i].ordinal));
                    SyncPage page;
                    B200_CUDA_TRY(cudaMemcpy(&page, ctx.d[i].window + 2ull * g * S, sizeof(page), cudaMemcpyDeviceToHost));
                    for (int t = 1; t < g; ++t) {
                        if (!page.start_ns[t] || page.done_ns[t] <= page.start_ns[t]) { stepped.clear(); break; }     // a launch that lost the barrier: shares
                        stepped[i][t].push_back((double)S / (double)(page.done_ns[t] - page.start_ns[t]));          // bytes per ns = GB/s
                    }
                }
            }
        }
        double mn = 1e300, mx = 0;
        for (int i = 0; i < g; ++i) {
            const double own = payload / (median_of(per_dev[i]) * 1e-3) / 1e9;      // the direction rank i's kernel drives
            const double common = payload / (out->ms_median * 1e-3) / 1e9;          // the other direction, over the common window
            out->egress_gbs[i] = mix ? own : pull ? common : own;       // MIX: rank i's kernel drives half of either direction
            out->ingress_gbs[i] = mix ? own : pull ? own : common;
            for (int j = 0; j < g; ++j) {
                if (i == j) continue;
                double gbs = own / (g - 1);                                         // per-pair share under full concurrency
                if (!stepped.empty()) gbs = median_of(stepped[i][(j - i + g) % g]);  // pair i -> j is step (j-i) mod G of rank i
                if (pair_gbs) pair_gbs[pull ? j * g + i : i * g + j] = gbs;
                mn = std::min(mn, gbs); mx = std::max(mx, gbs);
            }
        }
        out->min_pair_gbs = mn; out->max_pair_gbs = mx;
        out->pair_source = stepped.empty() ? B200PROBE_PAIR_SHARE : B200PROBE_PAIR_STEPPED;
    }
    rc = local_slots();
    if (rc) return rc;

    if (cfg.verify) {
        // every chunk that landed must equal the regenerated pattern under its (src,dst) seed
        int ok = 1;
        for (int dst = 0; dst < g && ok; ++dst) {
            PerDev& p = ctx.d[dst];
            B200_CUDA_TRY(cudaSetDevice(p.ordinal));
            for (int src = 0; src < g; ++src) {
                b200::VerifyResult v;
                rc = b200::verify_pattern(p.ordinal, p.window + (size_t)src * S, S, chunk_seed(cfg.seed, src, dst), p.partials, p.stream, &v);
                if (rc) return rc;
                if (v.bad) {
                    b200::set_error("a2a: chunk %d->%d landed with %llu wrong words (first at word %llu)", src, dst, (unsigned long long)v.bad,
                                    (unsigned long long)v.first);
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

int b200probe_nvlink_a2a(const int* ordinals, int g, const b200probe_a2a_cfg_t* cfg, double* pair_gbs, b200probe_a2a_result_t* out) {
    MutexLock lock(&g_ctx_mu);
    const int rc = nvlink_a2a_locked(ordinals, g, cfg, pair_gbs, out);
    if (rc && rc != B200PROBE_EMISMATCH) {      // a failed call leaves streams / windows in an unknown state: start over next time
        delete g_ctx;
        g_ctx = nullptr;
    }
    return rc;
}

}  // extern "C"
