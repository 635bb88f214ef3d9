// Shared helpers for the tvidz_b200 C-ABI library (sm_100a only).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tvidz_b200.h"

namespace tvz {

constexpr int kNumSMsB200 = 148;

char *err_buf();                       // thread-local, 512 bytes
int set_error(int code, const char *fmt, ...);
int num_sms();                          // SM count of the current device (cached per device)

// Fused gather (multi-GPU matcher): where the compaction's last block stores this rank's hit
// record on every peer, and the flag it raises there afterwards.
constexpr int kMaxPeers = 8;
struct GatherTargets {
    int n_peers = 0;                 // 0 = no gather
    unsigned epoch = 0;
    int *record[kMaxPeers] = {};     // this rank's slot inside peer p's gather buffer (peer memory)
    unsigned *flag[kMaxPeers] = {};  // this rank's flag word on peer p
};

// Batched compaction: blockIdx.y selects the query; per-query strides into the arrays.
struct BatchStrides {
    long long counts = 0, out = 0, rows = 0, state = 0;  // elements; ticket stride is 4 words, n_hits 1
};

// match.cu: single-pass ordered compaction shared by find_duplicates and fragment mode.
// `ticket` holds three u32 {next ticket = 0, query epoch = 1, finished blocks = 0}.
int compact_blocks(long long n_rows);
int compact_enqueue(int *counts, long long n_rows, int min_match, const int *vid, int *out, long long *rows_out,
                    long long cap, long long *n_hits_out, unsigned long long *state, unsigned *ticket,
                    const int *aux, int *aux_out, cudaStream_t st, const GatherTargets *gather = nullptr);
int compact_enqueue_keys(unsigned long long *keys, long long n_rows, int min_match, const int *vid, int *out,
                         long long *rows_out, long long cap, long long *n_hits_out, unsigned long long *state,
                         unsigned *ticket, int *delta_out, cudaStream_t st);
int gather_wait_enqueue(const unsigned *d_flags, int n_peers, unsigned epoch, cudaStream_t st);
int compact_enqueue_batch(int *counts, long long n_rows, int min_match, const int *vid, int *out, long long *rows_out,
                          long long cap, long long *n_hits_out, unsigned long long *state, unsigned *ticket, int n_batch,
                          const BatchStrides &bs, cudaStream_t st);

// Fragment mode: a row's best candidate as ONE u64 so that concurrent warps can combine theirs with
// atomicMax.  Order = the spec's: higher score, then smaller |d|, then smaller d.  |d| < 2^30.
__host__ __device__ __forceinline__ unsigned long long frag_key(int score, int d) {
    const unsigned ad = static_cast<unsigned>(d < 0 ? -d : d);
    return (static_cast<unsigned long long>(static_cast<unsigned>(score)) << 32) |
           (0xffffffffu - (2u * ad + (d > 0 ? 1u : 0u)));
}
__host__ __device__ __forceinline__ int frag_key_score(unsigned long long k) { return static_cast<int>(k >> 32); }
__host__ __device__ __forceinline__ int frag_key_delta(unsigned long long k) {
    if (k == 0) return 0;  // no candidate: (score 0, offset 0)
    const unsigned x = 0xffffffffu - static_cast<unsigned>(k);
    const int ad = static_cast<int>(x >> 1);
    return (x & 1u) ? ad : -ad;
}

}  // namespace tvz
#include <exception>
#include <new>
namespace tvz {
// The C ABI never throws: host-side allocation failures become TVZ_ERR_NOMEM.
template <class F>
int guarded(F &&body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc &) {
        return set_error(TVZ_ERR_NOMEM, "out of host memory");
    } catch (const std::exception &e) {
        return set_error(TVZ_ERR_INVALID, "unexpected exception: %s", e.what());
    } catch (...) {
        return set_error(TVZ_ERR_INVALID, "unexpected exception");
    }
}

#define TVZ_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return tvz::set_error(TVZ_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

#define TVZ_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) return tvz::set_error(TVZ_ERR_INVALID, __VA_ARGS__);                     \
    } while (0)

// ---- programmatic dependent launch (PDL) ----
// The kernels of one query (count -> compaction [-> next query's count]) are launched with the
// programmatic-stream-serialization attribute: a kernel's CTAs may become resident and run their
// prologue (first loads, shared-memory tables) while the predecessor's last CTAs drain, and block
// in pdl_wait() until the predecessor has completed and flushed -- which takes the grid launch
// latency off the critical path of a 50 us query.  Every such kernel calls pdl_wait() before it
// touches anything its predecessor writes (on all paths) and pdl_launch_dependents() right after,
// so a kernel two places down the stream never starts before the kernel two places up has finished.
bool pdl_enabled();  // common.cu: off with TVZ_NO_PDL=1
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// `cooperative`: all CTAs resident at once, so the kernel may use grid-wide barriers.
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool cooperative,
                       Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2]{};
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cooperative ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- mbarrier / TMA bulk-copy PTX wrappers (shared::cta addresses as u32) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
// try_wait suspends for a hardware-chosen interval, so 2^26 polls are many seconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
    unsigned polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++polls == (1u << 26)) {
#ifdef TVZ_DEBUG_WAIT
            printf("tvidz_b200: mbarrier wait timed out (tag %d, block %d, thread %d, bar 0x%x, parity %u)\n", tag,
                   (int)blockIdx.x, (int)threadIdx.x, bar, parity);
#endif
            (void)tag;
            __trap();
        }
    }
}
// TMA 1-D bulk copy global -> shared::cta, completion reported on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}


// ---- fused compaction (single launch for a whole query) ----
// After the last candidate has been resolved the grid synchronises (cooperative launch: all CTAs are
// resident), every CTA counts the qualifying rows of its chunks (kThreads * 8 rows each), the grid
// synchronises again, and every CTA writes its chunks' rows behind the hits of all earlier chunks --
// the same ordered record as match_compact_kernel, without a second launch, its cold start and the
// ticket / look-back protocol.  The per-row input (counts[], or the fragment kernel's packed keys)
// is zeroed on the way.
constexpr int kFusedRowsPerThread = 8;
struct FusedCompact {
    int enabled = 0;
    int min_match = 0;
    long long n_rows = 0, cap = 0;
    const int *vid = nullptr;
    int *out = nullptr;
    long long *rows_out = nullptr;
    long long *n_hits_out = nullptr;
    unsigned *chunk_hits = nullptr;      // [n_chunks]
    unsigned *done = nullptr;            // fused gather: CTAs finished
    unsigned long long *keys = nullptr;  // kKeys: per-row input
    int *aux_out = nullptr;              // kKeys: decoded offsets, aux_out[1 + hit]
    GatherTargets gt;
};

template <int kThreads, bool kKeys>
__device__ __forceinline__ void fused_compact(const FusedCompact &fc, int *__restrict__ counts, int *ws32) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    constexpr int kChunk = kThreads * kFusedRowsPerThread;
    constexpr int kWarps = kThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n_chunks = (fc.n_rows + kChunk - 1) / kChunk;
    __shared__ long long s_excl;
    // the thread's 8 rows: score / count in c[], (kKeys) decoded offset in d[]
    auto load8 = [&](long long r0, int (&c)[kFusedRowsPerThread], int (&d)[kFusedRowsPerThread]) {
        if (kKeys) {
#pragma unroll
            for (int j = 0; j < kFusedRowsPerThread; j += 2) {
                ulonglong2 t = make_ulonglong2(0ull, 0ull);
                if (r0 + j + 2 <= fc.n_rows) t = *reinterpret_cast<const ulonglong2 *>(fc.keys + r0 + j);
                else if (r0 + j < fc.n_rows) t.x = fc.keys[r0 + j];
                c[j] = frag_key_score(t.x); d[j] = frag_key_delta(t.x);
                c[j + 1] = frag_key_score(t.y); d[j + 1] = frag_key_delta(t.y);
            }
        } else if (r0 + kFusedRowsPerThread <= fc.n_rows) {
            const int4 a = *reinterpret_cast<const int4 *>(counts + r0);
            const int4 b = *reinterpret_cast<const int4 *>(counts + r0 + 4);
            c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < kFusedRowsPerThread; ++j) c[j] = r0 + j < fc.n_rows ? counts[r0 + j] : 0;
        }
    };
    grid.sync();  // every count / key of this query is final
    // pass A: hits per chunk
    for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const long long r0 = ch * kChunk + threadIdx.x * kFusedRowsPerThread;
        int c[kFusedRowsPerThread], d[kFusedRowsPerThread];
        load8(r0, c, d);
        int mine = 0;
#pragma unroll
        for (int j = 0; j < kFusedRowsPerThread; ++j) mine += c[j] >= fc.min_match && r0 + j < fc.n_rows;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        __syncthreads();
        if (lane == 0) ws32[warp] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int w = 0; w < kWarps; ++w) t += ws32[w];
            fc.chunk_hits[ch] = t;
        }
    }
    grid.sync();  // every chunk's hit count is published
    // pass B: ordered emission
    for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        __syncthreads();
        if (warp == 0) {  // hits of all earlier chunks
            long long e = 0;
            for (long long i = lane; i < ch; i += 32) e += fc.chunk_hits[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            if (lane == 0) s_excl = e;
        }
        const long long r0 = ch * kChunk + threadIdx.x * kFusedRowsPerThread;
        int c[kFusedRowsPerThread], d[kFusedRowsPerThread];
        load8(r0, c, d);
        int mine = 0;  // counts / scores are never negative: -1 marks "does not qualify / no such row"
#pragma unroll
        for (int j = 0; j < kFusedRowsPerThread; ++j) {
            if (r0 + j < fc.n_rows) {
                if (kKeys) { if (c[j] != 0 || d[j] != 0) fc.keys[r0 + j] = 0; }
                else if (c[j] != 0) counts[r0 + j] = 0;
                if (c[j] >= fc.min_match) ++mine; else c[j] = -1;
            } else {
                c[j] = -1;
            }
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) ws32[warp] = incl;
        __syncthreads();
        int wofs = 0;
        for (int w = 0; w < warp; ++w) wofs += ws32[w];
        long long pos = s_excl + wofs + (incl - mine);
#pragma unroll
        for (int j = 0; j < kFusedRowsPerThread; ++j) {
            if (c[j] >= 0) {
                if (pos < fc.cap) {
                    const int v = fc.vid[r0 + j];
                    fc.out[2 + 2 * pos] = v;
                    fc.out[3 + 2 * pos] = c[j];
                    fc.rows_out[pos] = r0 + j;
                    if (kKeys) fc.aux_out[1 + pos] = d[j];
                    for (int p2 = 0; p2 < fc.gt.n_peers; ++p2)
                        *reinterpret_cast<int2 *>(fc.gt.record[p2] + 2 + 2 * pos) = make_int2(v, c[j]);
                }
                ++pos;
            }
        }
        if (ch == n_chunks - 1 && threadIdx.x == kThreads - 1) {  // the last thread of the last chunk knows the total
            const long long total = pos;
            *fc.n_hits_out = total;
            fc.out[0] = total > 0x7fffffffll ? 0x7fffffff : static_cast<int>(total);
            fc.out[1] = total > fc.cap ? 1 : 0;
        }
    }
    if (fc.gt.n_peers == 0) return;
    // fused gather epilogue: the CTA that finishes last publishes the header and the flag on every peer
    __shared__ unsigned s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned dn = atomicAdd(fc.done, 1u);
        s_last = dn == gridDim.x - 1;
        if (s_last) *fc.done = 0;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < fc.gt.n_peers) {
        __threadfence_system();
        const int2 hdr = make_int2(*reinterpret_cast<volatile int *>(fc.out), *reinterpret_cast<volatile int *>(fc.out + 1));
        *reinterpret_cast<int2 *>(fc.gt.record[threadIdx.x]) = hdr;
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(fc.gt.flag[threadIdx.x]), "r"(fc.gt.epoch) : "memory");
    }
}

}  // namespace tvz
