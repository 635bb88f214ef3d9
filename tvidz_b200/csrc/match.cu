// Stage 2: find_duplicates over a device-resident packed catalogue.
//
// Replaces inspector/db.py:76-94 (full-table fetch at db.py:83 + the O(N*q*L) Python
// membership loop at db.py:85-91).  Semantics:
//     match_count(row) = #{ i : q[i] == some element of row }      (float ==)
//     result = [(video_id, match_count) for rows with match_count >= min_match]
//
// Data layout in HBM (one shard per GPU):
//     ts   : uint64 [n_vals]  IEEE-754 bit patterns of the stored timestamps, canonicalised
//            at pack time (-0.0 -> +0.0, NaN dropped, in-row repeats dropped) so that bitwise
//            equality == Python float equality and every stored value can add at most once
//     fp   : uint16 [n_vals]  filter_hash(ts[i]): what the single-query count kernel streams
//     off  : int64 [n_rows+1] CSR row offsets into ts
//     vid  : int32 [n_rows]   videos.id of each row
// Because in-row repeats are gone, match_count(row) = sum over stored values v of
// mult(v), where mult(v) = number of query positions equal to v.  The count kernel is
// therefore a pure streaming pass: 256-bit coalesced loads, a 64 KB shared-memory byte map of
// the query rejects ~all values with one LDS.U8, and the rare survivors are looked up exactly
// and added to counts[row] with a RED.  One query streams the 2-byte fingerprints (`fp`) and
// touches `ts` only for survivors; the batched kernel (8 queries per pass) streams `ts` itself.
#include <algorithm>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace tvz {
namespace {

constexpr int kMapEntries = 1 << 16;         // byte-map filter of the query: 64 KB of shared memory
#ifndef TVZ_MAX_KEYS
#define TVZ_MAX_KEYS 2048
#endif
constexpr int kMaxKeys = TVZ_MAX_KEYS;       // distinct query values per launch
constexpr int kCountThreads = 512;            // batched kernel (streams the 8-byte values)
constexpr int kCountUnroll = 8;               // 8 x 16 B (= 4 x 256-bit loads) in flight per thread
constexpr int kChunkPairs = kCountThreads * kCountUnroll;  // 4096 pairs = 8192 values per CTA iteration
constexpr int kCountWarps = kCountThreads / 32;
constexpr int kWarpQueue = 64;               // filter survivors parked per warp
constexpr int kBlockShift = 7;               // coarse row index: one entry per 128 stored values
// 16384 rows per block: the look-back walks its predecessors 32 at a time, and every hop is a dependent
// global round trip -- 1 M rows are 62 blocks (<= 2 hops) instead of 489 (<= 15 hops, ~8 us)
constexpr int kScanThreads = 1024;
constexpr int kScanRowsPerThread = 16;
constexpr int kScanRowsPerBlock = kScanThreads * kScanRowsPerThread;
constexpr unsigned long long kPadPattern = 0x7ff8dead0000beefull;  // a NaN: never equals a stored value

// Two IMADs and a shift: good enough on frame-quantised timestamps and on x.0 / x.5 values
// (false-positive rate ~ n_keys / 65536, measured in DESIGN.md).
__host__ __device__ __forceinline__ uint32_t filter_hash(unsigned long long v) {
    const uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
    return (lo * 0x9E3779B1u + hi * 0x85EBCA77u) >> 16;
}

// counts[row] += mult(v) for every stored value v that equals a query value.
//
// Streaming pass over `ts` (padded to whole chunks, so the hot loop has no bounds checks):
// per value one hash, one LDS.U8 and a warp vote; the next chunk's 128-bit loads are already
// in flight while the current one is probed.  Survivors (true matches plus ~0.1% false
// positives) are parked in a per-warp shared-memory queue -- slots handed out from the vote
// mask, no atomics -- and resolved 32 at a time by the whole warp: exact key lookup, then the
// row through a coarse index (row of every 256th value) and a short search in `off`.  A hit
// therefore never stalls the other 31 lanes, and no CTA-wide barrier sits in the loop.
// Short queries (the common case: a video has tens of cuts) ride in the kernel parameters,
// which saves the two host->device copies in front of the launch.
constexpr int kParamKeys = 224;
struct SmallQuery {
    unsigned long long keys[kParamKeys];
    int mult[kParamKeys];
};

// 256-bit streaming load (sm_100 LDG.E.256): no L1 allocation -- this kernel leaves L1 almost
// no room, shared memory takes ~213 of the SM's 228 KB -- and evict-first in L2.
struct U64x4 {
    unsigned long long a, b, c, d;
};
__device__ __forceinline__ U64x4 ld_stream_256(const void *p) {
    U64x4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d)
                 : "l"(p));
    return r;
}

// Which 32-byte unit of the CTA's chunk thread t loads as its j-th (batched kernel).
__device__ __forceinline__ int unit_index(int j) { return j * kCountThreads + threadIdx.x; }

// ---- single query: streaming pass over the 16-bit FINGERPRINTS of the stored values ----
//
// The byte-map filter only ever looks at filter_hash(v), 16 bits of a value.  The catalogue
// therefore also stores fp[i] = filter_hash(ts[i]) as a uint16 array (+2 B per stored value), and the
// count kernel streams THAT: 2 bytes of HBM traffic per stored timestamp instead of 8 -- 128 MB
// instead of 512 MB for the 1 M-row catalogue.  Nothing is lost: a fingerprint that passes the
// filter (true matches + ~n_keys/65536 false positives, ~0.3 % of values) is verified against the
// real 8-byte value, fetched from `ts` only then, with the exact key lookup as before.
//
// A warp-wide 256-bit load brings 512 fingerprints (16 per lane).  Units are dealt round-robin to
// all warps of the grid (unit g = (it * kFpUnits + j) * n_warps + warp), so the grid sweeps one
// contiguous window per load slot and every warp ends within one unit of every other.  Per
// fingerprint: extract, LDS.U8, shift-or into the lane's 16-bit survivor mask; one vote per unit.
// Survivors are parked inline (slots from a shuffle scan) as element indices in a per-warp queue and
// resolved 32 at a time -- normally once, when the warp has finished streaming: load the value,
// binary search of the sorted query keys, row through the coarse index + a short search in `off`,
// RED.ADD counts[row] += mult.
// launch shape (overridable for tuning sweeps: scripts/sweep_match.py builds variants)
#ifndef TVZ_FP_THREADS
#define TVZ_FP_THREADS 512
#endif
#ifndef TVZ_FP_UNITS
#define TVZ_FP_UNITS 2   // measured: 2 -> 48.7 us, 4 -> 59.4 us (spills at 64 registers), 8 at 256 threads -> 53.2 us
#endif
constexpr int kFpThreads = TVZ_FP_THREADS;
constexpr int kFpWarps = kFpThreads / 32;
constexpr int kFpUnits = TVZ_FP_UNITS;       // 256-bit loads in flight per thread
constexpr int kFpPerUnit = 32 * 16;          // fingerprints per warp-wide load
#ifndef TVZ_FP_QUEUE
#define TVZ_FP_QUEUE 128
#endif
#ifndef TVZ_FP_MINB
#define TVZ_FP_MINB (TVZ_FP_THREADS >= 1024 ? 1 : 2)
#endif
constexpr int kFpQueue = TVZ_FP_QUEUE;       // survivors parked per warp

struct alignas(16) FpSmem {
    unsigned char map[kMapEntries];   // first: zeroed with 16-byte stores
    unsigned long long keys[kMaxKeys];
    long long qe[kFpWarps][kFpQueue];
    int mult[kMaxKeys];
};

struct U32x8 {
    unsigned w[8];
};
__device__ __forceinline__ U32x8 ld_stream_u32x8(const void *p) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
    U32x8 r;
    r.w[0] = static_cast<unsigned>(a); r.w[1] = static_cast<unsigned>(a >> 32);
    r.w[2] = static_cast<unsigned>(b); r.w[3] = static_cast<unsigned>(b >> 32);
    r.w[4] = static_cast<unsigned>(c); r.w[5] = static_cast<unsigned>(c >> 32);
    r.w[6] = static_cast<unsigned>(d); r.w[7] = static_cast<unsigned>(d >> 32);
    return r;
}

// What a surviving fingerprint is checked against: the value behind it and the row it belongs to, one
// 16-byte record per arranged position (one DRAM fetch per survivor).
struct alignas(16) VerifyRec {
    unsigned long long ts;
    unsigned row, pad;
};
struct FpCtx {
    const VerifyRec *rec;
    int *counts;
    int n_keys;
};

// Verify one survivor: ONE memory round trip (value and row of the arranged position, fetched together
// and already on their way to L2 since the survivor was parked), then the sorted query keys are
// searched in shared memory and a real match adds into counts[row].
__device__ __forceinline__ void fp_resolve(const FpCtx &cx, const FpSmem &sm, long long pos) {
    const uint4 r4 = __ldg(reinterpret_cast<const uint4 *>(cx.rec + pos));
    const unsigned long long v = (static_cast<unsigned long long>(r4.y) << 32) | r4.x;
    const unsigned row = r4.z;
    int lo = 0, hi = cx.n_keys;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sm.keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    if (lo >= cx.n_keys || sm.keys[lo] != v) return;  // fingerprint collision (or padding)
    atomicAdd(&cx.counts[row], sm.mult[lo]);
}

// Verify the warp's parked survivors, 32 at a time (called by the whole warp).  Out of line for the
// rare mid-stream call (a dense query filling the queue), inline for the one at the end of the stream.
__device__ __forceinline__ void fp_drain_inline(const FpCtx &cx, const FpSmem &sm, const long long *qe, int n) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    for (int i = lane; i < n; i += 32) fp_resolve(cx, sm, qe[i]);
    __syncwarp();
}
__device__ __noinline__ void fp_drain(const FpCtx &cx, const FpSmem &sm, const long long *qe, int n) {
    fp_drain_inline(cx, sm, qe, n);
}

constexpr int kFusedChunk = kFpThreads * kFusedRowsPerThread;   // rows per chunk of the count kernel's fused compaction

template <bool kParamQuery>
__global__ void __launch_bounds__(kFpThreads, TVZ_FP_MINB)
match_count_kernel(const unsigned short *__restrict__ fp, long long n_units,
                   const VerifyRec *__restrict__ rec,
                   const unsigned long long *__restrict__ keys, const int *__restrict__ mult, int n_keys,
                   int *__restrict__ counts, const __grid_constant__ SmallQuery sq,
                   const __grid_constant__ FusedCompact fc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FpSmem &sm = *reinterpret_cast<FpSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the first loads do not depend on the query: issue them before the byte map is built
    const long long n_warps = static_cast<long long>(gridDim.x) * kFpWarps;
    const long long wg = static_cast<long long>(blockIdx.x) * kFpWarps + warp;
    const unsigned short *lane_fp = fp + lane * 16;
    U32x8 v[kFpUnits];
#pragma unroll
    for (int j = 0; j < kFpUnits; ++j) {
        const long long g = j * n_warps + wg;
        if (g < n_units) v[j] = ld_stream_u32x8(lane_fp + g * kFpPerUnit);
    }
    for (int i = threadIdx.x; i < kMapEntries / 16; i += kFpThreads)
        reinterpret_cast<uint4 *>(sm.map)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int i = threadIdx.x; i < n_keys; i += kFpThreads) {
        const unsigned long long k = kParamQuery ? sq.keys[i] : keys[i];
        sm.keys[i] = k;
        sm.mult[i] = kParamQuery ? sq.mult[i] : mult[i];
        sm.map[filter_hash(k)] = 1;
    }
    __syncthreads();
    pdl_wait();  // counts[] is being read and zeroed by the previous query's compaction until here
    pdl_launch_dependents();  // only now: this query's compaction takes its tickets before ITS wait

    const FpCtx cx{rec, counts, n_keys};
    long long *qe = sm.qe[warp];
    int queued = 0;  // warp-uniform, <= kFpQueue

    for (long long g0 = wg; g0 < n_units; g0 += kFpUnits * n_warps) {
#pragma unroll
        for (int j = 0; j < kFpUnits; ++j) {
            const long long g = g0 + j * n_warps;
            if (g >= n_units) break;  // warp-uniform
            unsigned flags = 0;       // bit k: the lane's k-th fingerprint is in the query's byte map
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned w = v[j].w[k];
#ifdef TVZ_FP_NOLOOKUP   // tuning: streaming floor without the byte-map lookups
                flags |= (w == 0x12345678u) << k;
#else
                flags |= static_cast<unsigned>(sm.map[w & 0xffffu]) << (2 * k);
                flags |= static_cast<unsigned>(sm.map[w >> 16]) << (2 * k + 1);
#endif
            }
#ifdef TVZ_FP_NOPARK     // tuning: lookups only, survivors dropped
            if (flags == 0xdeadbeefu) qe[0] = 1;
            flags = 0;
#endif
            // park the survivors' element indices, one per lane and round (a lane rarely holds two):
            // slots from the vote mask, no atomics, no scan
            unsigned mask = __ballot_sync(0xffffffffu, flags != 0u);
            if (mask) {
                const long long e0 = g * kFpPerUnit + lane * 16;
                do {
                    if (queued + __popc(mask) > kFpQueue) {
                        fp_drain(cx, sm, qe, queued);
                        queued = 0;
                    }
                    if (flags) {
                        const long long e = e0 + (__ffs(flags) - 1);
                        qe[queued + __popc(mask & ((1u << lane) - 1u))] = e;
                        flags &= flags - 1;
                        // what the verification will read, on its way to L2 while the stream goes on
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + e));
                    }
                    queued += __popc(mask);
                    mask = __ballot_sync(0xffffffffu, flags != 0u);
                } while (mask);
            }
            const long long gn = g + kFpUnits * n_warps;
            if (gn < n_units) v[j] = ld_stream_u32x8(lane_fp + gn * kFpPerUnit);
        }
    }
    fp_drain_inline(cx, sm, qe, queued);
    if (fc.enabled) fused_compact<kFpThreads, false>(fc, counts, sm.mult);  // (the mult table is dead by now: scratch)
}

// ---- batched queries: up to 8 find_duplicates calls answered by ONE pass over the catalogue ----
// (the reference runs one analysis thread per upload, app.py:43,472, each calling find_duplicates:
// concurrent queries are the normal case).  The byte map holds one bit per query, so one LDS.U8
// says which of the 8 queries might contain a value; survivors carry that mask through the
// per-warp queue and add into counts[query][row].
constexpr int kBatch = 8;
struct alignas(16) BatchSmem {
    unsigned char map[kMapEntries];
    unsigned long long keys[kBatch][kParamKeys];
    unsigned long long qv[kCountWarps][kWarpQueue];
    long long qe[kCountWarps][kWarpQueue];            // value index | query mask << 56
    int mult[kBatch][kParamKeys];
    int n_keys[kBatch];
};

__global__ void __launch_bounds__(kCountThreads, 2)
match_count_batch_kernel(const ulonglong2 *__restrict__ ts2, long long n_pairs_padded,
                         const unsigned long long *__restrict__ keys, const int *__restrict__ mult,
                         const int *__restrict__ n_keys, int n_batch, const long long *__restrict__ off,
                         const int *__restrict__ block_row, long long n_rows, int *__restrict__ counts,
                         long long counts_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BatchSmem &sm = *reinterpret_cast<BatchSmem *>(smem_raw);
    pdl_wait();  // keys / n_keys are uploaded, counts[] zeroed by what runs before
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < kMapEntries / 16; i += kCountThreads)
        reinterpret_cast<uint4 *>(sm.map)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < kBatch) sm.n_keys[threadIdx.x] = threadIdx.x < n_batch ? n_keys[threadIdx.x] : 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_batch * kParamKeys; i += kCountThreads) {
        const int b = i / kParamKeys, k = i - b * kParamKeys;
        if (k < sm.n_keys[b]) {
            const unsigned long long key = keys[i];
            sm.keys[b][k] = key;
            sm.mult[b][k] = mult[i];
            // byte-wide atomic OR through the containing 32-bit word
            const uint32_t h = filter_hash(key);
            atomicOr(reinterpret_cast<unsigned *>(sm.map) + (h >> 2), (1u << b) << (8 * (h & 3)));
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    unsigned long long *qv = sm.qv[threadIdx.x >> 5];
    long long *qe = sm.qe[threadIdx.x >> 5];
    int queued = 0;  // warp-uniform

    auto resolve = [&](unsigned long long v, long long packed) {
        const long long elem = packed & ((1ll << 56) - 1);
        unsigned qmask = static_cast<unsigned>(static_cast<unsigned long long>(packed) >> 56);
        long long row = -1;
        while (qmask) {
            const int b = __ffs(qmask) - 1;
            qmask &= qmask - 1;
            int lo = 0, hi = sm.n_keys[b];
            const int nk = hi;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sm.keys[b][mid] < v) lo = mid + 1; else hi = mid;
            }
            if (lo >= nk || sm.keys[b][lo] != v) continue;  // filter false positive for this query
            if (row < 0) {
                const long long blk = elem >> kBlockShift;
                long long a = block_row[blk], z = block_row[blk + 1] + 1;
                while (z - a > 1) {
                    const long long mid = (a + z) >> 1;
                    if (off[mid] <= elem) a = mid; else z = mid;
                }
                row = a;
            }
            atomicAdd(&counts[b * counts_stride + row], sm.mult[b][lo]);
        }
    };
    auto drain = [&]() {
        __syncwarp();
        for (int i = lane; i < min(queued, kWarpQueue); i += 32) resolve(qv[i], qe[i]);
        __syncwarp();
        queued = 0;
    };
    auto park = [&](unsigned m, unsigned long long x, long long elem) {
        const unsigned mask = __ballot_sync(0xffffffffu, m != 0);
        if (m) {
            const int slot = queued + __popc(mask & ((1u << lane) - 1u));
            const long long packed = elem | (static_cast<long long>(m) << 56);
            if (slot < kWarpQueue) { qv[slot] = x; qe[slot] = packed; }
            else resolve(x, packed);
        }
        queued += __popc(mask);
    };

    constexpr int kUnits = kCountUnroll / 2;
    const long long stride = static_cast<long long>(gridDim.x) * kChunkPairs;
    long long base = static_cast<long long>(blockIdx.x) * kChunkPairs;
    const U64x4 *ts4 = reinterpret_cast<const U64x4 *>(ts2);
    U64x4 v[kUnits];
    if (base < n_pairs_padded) {
#pragma unroll
        for (int j = 0; j < kUnits; ++j) v[j] = ld_stream_256(ts4 + (base >> 1) + unit_index(j));
    }
    for (; base < n_pairs_padded; base += stride) {
        const bool more = base + stride < n_pairs_padded;
        const U64x4 *next = ts4 + ((base + stride) >> 1);
#pragma unroll
        for (int j = 0; j < kUnits; ++j) {
            const unsigned ma = sm.map[filter_hash(v[j].a)], mb = sm.map[filter_hash(v[j].b)];
            const unsigned mc = sm.map[filter_hash(v[j].c)], md = sm.map[filter_hash(v[j].d)];
            const unsigned any = __reduce_or_sync(0xffffffffu, (ma != 0) | ((mb != 0) << 1) | ((mc != 0) << 2) |
                                                                   ((md != 0) << 3));
            if (any) {
                const long long elem = 2 * base + 4 * unit_index(j);
                if (any & 1u) park(ma, v[j].a, elem);
                if (any & 2u) park(mb, v[j].b, elem + 1);
                if (any & 4u) park(mc, v[j].c, elem + 2);
                if (any & 8u) park(md, v[j].d, elem + 3);
                if (queued >= kWarpQueue / 2) drain();
            }
            if (more) v[j] = ld_stream_256(next + unit_index(j));
        }
    }
    drain();
}

// Ordered compaction of the rows with counts[row] >= min_match, in ONE pass (decoupled
// look-back): a block takes a ticket (so tickets start in order), counts its qualifying rows,
// publishes {epoch, AGGREGATE, n}, sums its predecessors' records walking backwards 32 at a
// time until it meets an inclusive PREFIX, publishes its own PREFIX, and writes its rows at
// that offset in row order; counts[] is zeroed for the next query.  Records carry the query
// epoch, so `state` never needs clearing.  out: int32 [cap+1][2]; out[0] = {n_hits saturated,
// overflow flag}; out[1+h] = {video_id, match_count}; rows_out[h] = row index.
constexpr unsigned long long kStateAggregate = 1ull << 32, kStatePrefix = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// kKeys (fragment mode, streaming kernel): the per-row input is a packed u64 best-candidate key
// (common.cuh: frag_key) instead of counts[] + aux[]; score = key >> 32, the offset is decoded.
template <bool kKeys>
__global__ void __launch_bounds__(kScanThreads)
match_compact_kernel(int *__restrict__ counts, long long n_rows, int min_match, const int *__restrict__ vid,
                     int *__restrict__ out, long long *__restrict__ rows_out, long long cap,
                     long long *__restrict__ n_hits_out, unsigned long long *state, unsigned *ticket,
                     const int *__restrict__ aux, int *__restrict__ aux_out, const __grid_constant__ GatherTargets gt,
                     const BatchStrides bs, unsigned long long *__restrict__ keys) {
    // batched queries: blockIdx.y picks the query, everything below is per query
    counts += blockIdx.y * bs.counts;
    out += blockIdx.y * bs.out;
    rows_out += blockIdx.y * bs.rows;
    state += blockIdx.y * bs.state;
    ticket += blockIdx.y * 4;
    n_hits_out += blockIdx.y;
    // ticket[0] = next ticket, ticket[1] = query epoch.  The epoch is read BEFORE the ticket is
    // taken and bumped by the holder of the last ticket, i.e. after every block has read it:
    // the kernel is self-contained and can be replayed from a CUDA graph.
    __shared__ unsigned s_block, s_epoch;
    __shared__ long long s_excl;
    __shared__ int ws[kScanThreads / 32];
    // Tickets are taken BEFORE the dependency wait: ticket and epoch are only ever touched by compaction
    // kernels, and the previous one has completed (the kernel in between waited for it before it let
    // this one launch), so these two global round trips hide behind the count kernel's tail.
    if (threadIdx.x == 0) {
        unsigned e;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(ticket + 1) : "memory");
        s_epoch = e;
        s_block = atomicAdd(ticket, 1u);
    }
    pdl_launch_dependents();
    pdl_wait();  // the count / fragment kernel before this one must have finished
    __syncthreads();
    const unsigned blk = s_block;
    const unsigned epoch = s_epoch;
    const unsigned long long tag = static_cast<unsigned long long>(epoch) << 34;
    const long long r0 = blk * static_cast<long long>(kScanRowsPerBlock) + threadIdx.x * kScanRowsPerThread;
    int cnt[kScanRowsPerThread];
    int dec[kKeys ? kScanRowsPerThread : 1];
    int mine = 0;
    // the thread's 16 rows in 128-bit loads (the arrays are cudaMalloc-aligned, r0 is a multiple of 16);
    // the last, partial thread range of a shard goes row by row
    const bool whole = r0 + kScanRowsPerThread <= n_rows;
    int cin[kKeys ? 1 : kScanRowsPerThread];
    unsigned long long kin[kKeys ? kScanRowsPerThread : 1];
    if (whole) {
        if (kKeys) {
#pragma unroll
            for (int j = 0; j < kScanRowsPerThread; j += 2) {
                const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(keys + r0 + j);
                kin[j] = t.x;
                kin[j + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kScanRowsPerThread; j += 4) {
                const int4 t = *reinterpret_cast<const int4 *>(counts + r0 + j);
                cin[j] = t.x; cin[j + 1] = t.y; cin[j + 2] = t.z; cin[j + 3] = t.w;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kScanRowsPerThread; ++j) {
        cnt[j] = -1;
        if (whole || r0 + j < n_rows) {
            int c;
            if (kKeys) {
                const unsigned long long k = whole ? kin[j] : keys[r0 + j];
                if (k != 0) keys[r0 + j] = 0;
                c = frag_key_score(k);
                dec[j] = frag_key_delta(k);
            } else {
                c = whole ? cin[j] : counts[r0 + j];
                if (c != 0) counts[r0 + j] = 0;
            }
            if (c >= min_match) { cnt[j] = c; ++mine; }
        }
    }
    // block-wide exclusive scan of `mine` (rows are thread-contiguous: thread order = row order)
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, d);
        if ((threadIdx.x & 31) >= d) incl += n;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int w = lane < kScanThreads / 32 ? ws[lane] : 0;
        int run = w;
#pragma unroll
        for (int d = 1; d < kScanThreads / 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, run, d);
            if (lane >= d) run += n;
        }
        if (lane < kScanThreads / 32) ws[lane] = run - w;  // exclusive warp offsets
        const unsigned agg = static_cast<unsigned>(__shfl_sync(0xffffffffu, run, kScanThreads / 32 - 1));
        long long excl = 0;
        if (blk == 0) {
            if (lane == 0) st_release_u64(&state[0], tag | kStatePrefix | agg);
        } else {
            if (lane == 0) st_release_u64(&state[blk], tag | kStateAggregate | agg);
            long long idx = static_cast<long long>(blk) - 1;
            while (true) {
                const long long i = idx - lane;
                unsigned long long rec = 0;
                unsigned prefix_mask, valid_mask;
                do {  // poll until the window up to the first PREFIX is published for this epoch
                    rec = i >= 0 ? ld_acquire_u64(&state[i]) : (tag | kStatePrefix);
                    const bool ok = (rec >> 34) == epoch && ((rec >> 32) & 3ull) != 0;
                    valid_mask = __ballot_sync(0xffffffffu, ok);
                    prefix_mask = __ballot_sync(0xffffffffu, ok && ((rec >> 32) & 3ull) == 2ull);
                    // lanes below the first PREFIX lane must all be valid
                } while ((prefix_mask ? ((valid_mask | ~((prefix_mask & -prefix_mask) - 1u)) != 0xffffffffu)
                                      : (valid_mask != 0xffffffffu)));
                const unsigned upto = prefix_mask ? (prefix_mask & -prefix_mask) : 0u;
                const unsigned take = prefix_mask ? ((upto - 1u) | upto) : 0xffffffffu;  // lanes 0..first PREFIX
                long long v = ((take >> lane) & 1u) ? static_cast<long long>(rec & 0xffffffffull) : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                excl += v;
                if (prefix_mask) break;
                idx -= 32;
            }
            // hits are bounded by rows < 2^32 per shard, so the running prefix fits 32 bits
            if (lane == 0) st_release_u64(&state[blk], tag | kStatePrefix | static_cast<unsigned>(excl + agg));
        }
        if (lane == 0) {
            s_excl = excl;
            if (blk == gridDim.x - 1) {  // last ticket: every block has its ticket, totals are final
                const long long total = excl + agg;
                *n_hits_out = total;
                out[0] = total > 0x7fffffffll ? 0x7fffffff : static_cast<int>(total);
                out[1] = total > cap ? 1 : 0;
                ticket[1] = (epoch + 1u) & 0x3fffffffu;  // every record is rewritten per query: no stale match
                __threadfence();
                ticket[0] = 0;
            }
        }
    }
    __syncthreads();
    long long pos = s_excl + ws[threadIdx.x >> 5] + (incl - mine);
#pragma unroll
    for (int j = 0; j < kScanRowsPerThread; ++j) {
        if (cnt[j] >= 0) {
            if (pos < cap) {
                out[2 + 2 * pos] = vid[r0 + j];
                out[3 + 2 * pos] = cnt[j];
                rows_out[pos] = r0 + j;
                // per-row payload (fragment mode: best offset)
                if (kKeys) aux_out[1 + pos] = dec[j];
                else if (aux) aux_out[1 + pos] = aux[r0 + j];
                // fused gather: every block ships its own hits to all peers (8-byte stores over NVLink)
                for (int p = 0; p < gt.n_peers; ++p)
                    *reinterpret_cast<int2 *>(gt.record[p] + 2 + 2 * pos) = make_int2(vid[r0 + j], cnt[j]);
            }
            ++pos;
        }
    }
    if (gt.n_peers == 0) return;

    // ---- fused gather epilogue: the block that finishes last publishes the header and the flag ----
    __shared__ unsigned s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();  // this block's peer stores (and the local header) before it counts as done
        const unsigned done = atomicAdd(ticket + 2, 1u);
        s_last = done == gridDim.x - 1;
        if (s_last) ticket[2] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < gt.n_peers) {
        __threadfence_system();
        const int2 hdr = make_int2(*reinterpret_cast<volatile int *>(out), *reinterpret_cast<volatile int *>(out + 1));
        *reinterpret_cast<int2 *>(gt.record[threadIdx.x]) = hdr;  // {n_hits, overflow}
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(gt.flag[threadIdx.x]), "r"(gt.epoch) : "memory");
    }
}

// Wait until every peer's record for `epoch` has landed in this rank's gather buffer.  Bounded:
// a peer that never answers turns into a launch failure, not a hung GPU.
__global__ void gather_wait_kernel(const unsigned *flags, int n_peers, unsigned epoch) {
    if (threadIdx.x >= n_peers) return;
    unsigned v, polls = 0;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
        if (v == epoch) return;
        __nanosleep(64);
    } while (++polls < (1u << 26));
    __trap();
}

// Per hit row: the 1-based query index at which the row reaches min_match.
// One warp per hit; q_canon holds the canonicalised query in query order (NaNs left as NaN
// bit patterns, which equal no stored value).
__global__ void match_kth_kernel(const unsigned long long *__restrict__ ts, const long long *__restrict__ off,
                                 const long long *__restrict__ rows, const int *__restrict__ out_hdr, long long cap,
                                 const unsigned long long *__restrict__ q_canon, int qn, int min_match,
                                 int *__restrict__ kth) {
    const long long n_hits = min(static_cast<long long>(out_hdr[0]), cap);
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long h = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; h < n_hits; h += warps) {
        const long long r = rows[h];
        const long long b = off[r], e = off[r + 1];
        int c = 0, k = 0;
        if (min_match >= 1) {
            for (int i = 0; i < qn && k == 0; ++i) {
                const unsigned long long v = q_canon[i];
                bool found = false;
                for (long long j = b + lane; j < e + ((32 - ((e - b) & 31)) & 31); j += 32) {
                    const bool m = j < e && ts[j] == v;
                    if (__any_sync(0xffffffffu, m)) { found = true; break; }
                }
                if (found && ++c >= min_match) k = i + 1;
            }
        }
        if (lane == 0) kth[h] = k;
    }
}

}  // namespace

int compact_blocks(long long n_rows) {
    return static_cast<int>((n_rows + kScanRowsPerBlock - 1) / kScanRowsPerBlock);
}

// Ordered compaction of counts[row] >= min_match (see match_compact_kernel).  `state` holds
// compact_blocks(n_rows) u64 records (zero-initialised once), `ticket` two u32 {0, 1}.
int compact_enqueue(int *counts, long long n_rows, int min_match, const int *vid, int *out, long long *rows_out,
                    long long cap, long long *n_hits_out, unsigned long long *state, unsigned *ticket,
                    const int *aux, int *aux_out, cudaStream_t st, const GatherTargets *gather) {
    const GatherTargets none{};
    TVZ_CUDA(launch_pdl(match_compact_kernel<false>, dim3(compact_blocks(n_rows)), dim3(kScanThreads), 0, st, false, counts,
                        n_rows, min_match, vid, out, rows_out, cap, n_hits_out, state, ticket, aux, aux_out,
                        gather ? *gather : none, BatchStrides{}, static_cast<unsigned long long *>(nullptr)));
    return TVZ_OK;
}

// Same compaction over packed best-candidate keys (fragment streaming kernel); keys[] is zeroed.
int compact_enqueue_keys(unsigned long long *keys, long long n_rows, int min_match, const int *vid, int *out,
                         long long *rows_out, long long cap, long long *n_hits_out, unsigned long long *state,
                         unsigned *ticket, int *delta_out, cudaStream_t st) {
    TVZ_CUDA(launch_pdl(match_compact_kernel<true>, dim3(compact_blocks(n_rows)), dim3(kScanThreads), 0, st, false,
                        static_cast<int *>(nullptr), n_rows, min_match, vid, out, rows_out, cap, n_hits_out, state, ticket,
                        static_cast<const int *>(nullptr), delta_out, GatherTargets{}, BatchStrides{}, keys));
    return TVZ_OK;
}

int compact_enqueue_batch(int *counts, long long n_rows, int min_match, const int *vid, int *out, long long *rows_out,
                          long long cap, long long *n_hits_out, unsigned long long *state, unsigned *ticket, int n_batch,
                          const BatchStrides &bs, cudaStream_t st) {
    const GatherTargets none{};
    dim3 grid(compact_blocks(n_rows), n_batch);
    TVZ_CUDA(launch_pdl(match_compact_kernel<false>, grid, dim3(kScanThreads), 0, st, false, counts, n_rows, min_match, vid, out,
                        rows_out, cap, n_hits_out, state, ticket, static_cast<const int *>(nullptr),
                        static_cast<int *>(nullptr), none, bs, static_cast<unsigned long long *>(nullptr)));
    return TVZ_OK;
}

int gather_wait_enqueue(const unsigned *d_flags, int n_peers, unsigned epoch, cudaStream_t st) {
    gather_wait_kernel<<<1, 32, 0, st>>>(d_flags, n_peers, epoch);
    TVZ_CUDA(cudaGetLastError());
    return TVZ_OK;
}

}  // namespace tvz

using namespace tvz;

struct tvz_catalog {
    int device = 0;
    long long n_rows = 0, n_vals = 0, n_pairs = 0;
    long long n_pairs_padded = 0;  // ts is padded with a NaN pattern to whole count-kernel chunks
    unsigned long long *d_ts = nullptr;
    unsigned short *d_fp = nullptr;  // filter_hash of every stored value, padded to whole 512-value units and
                                     // arranged inside each unit for conflict-free lookups (arrange_fingerprints)
    void *d_rec = nullptr;           // VerifyRec[p]: the value behind d_fp[p] and its row, in the same arranged order
    long long n_units = 0;
    long long *d_off = nullptr;
    int *d_vid = nullptr;
    int *d_block_row = nullptr;    // row holding stored value b*1024 (coarse index for hit -> row)
};

struct tvz_match_ws {
    const tvz_catalog *cat = nullptr;
    long long cap = 0;
    int n_blocks = 0;
    int *d_counts = nullptr;       // [n_rows], zero between queries
    unsigned long long *d_state = nullptr;  // [n_blocks] look-back records {epoch, flag, value}
    unsigned *d_ticket = nullptr;           // {next ticket, query epoch}
    int *d_out = nullptr;          // [cap+1][2]
    long long *d_rows = nullptr;   // [cap]
    int *d_kth = nullptr;          // [cap]
    long long *d_nhits = nullptr;  // [1]
    unsigned *d_chunk_hits = nullptr;  // [n_rows / 4096 + 1] fused compaction: qualifying rows per chunk
    // query staging: keys u64 [kMaxKeys] | q_canon u64 [q_cap] | mult i32 [kMaxKeys]
    unsigned long long *d_keys = nullptr, *d_qcanon = nullptr;
    int *d_mult = nullptr;
    uint8_t *h_stage = nullptr;    // pinned
    size_t stage_bytes = 0;
    int q_cap = 0;
    int *h_out = nullptr;          // pinned [cap+1][2]
    int *h_kth = nullptr;          // pinned [cap]
    cudaStream_t stream = nullptr; // private stream for the synchronous entry point
    cudaEvent_t staged = nullptr;  // the pinned staging buffer has been consumed
    bool stage_busy = false;
    bool timing = false;           // debug: bracket the count kernel(s) with events
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    // batched queries (allocated on first use): per-query copies of counts/out/rows/state/ticket
    int *b_counts = nullptr, *b_out = nullptr, *b_mult = nullptr, *b_nkeys = nullptr;
    long long *b_rows = nullptr, *b_nhits = nullptr;
    unsigned long long *b_state = nullptr, *b_keys = nullptr;
    unsigned *b_ticket = nullptr;
    uint8_t *hb_stage = nullptr;   // pinned: keys | mult | n_keys
    int *hb_out = nullptr;         // pinned [8][cap+1][2]
    long long hb_cap = 0;
};

namespace {

inline unsigned long long canon_bits(double v) {
    unsigned long long b;
    memcpy(&b, &v, 8);
    if (b == 0x8000000000000000ull) b = 0;  // -0.0 == 0.0 in Python
    return b;
}
inline bool is_nan_bits(unsigned long long b) {
    return (b & 0x7ff0000000000000ull) == 0x7ff0000000000000ull && (b & 0x000fffffffffffffull) != 0;
}

// Fingerprints of one 512-value unit, ARRANGED so that the k-th lookups of the 32 lanes (positions
// lane * 16 + k, k = 0..15: one "round") fall into 32 different shared-memory banks wherever the unit
// allows it.  The byte-map lookup of fingerprint f goes to bank (f >> 2) & 31 whatever the query is,
// so this is decided once, when the catalogue is packed.  A bank with exactly 16 of the unit's 512
// values appears once in every round.  Banks with more (their extras) and banks with fewer (their
// absences) are confined to the LAST T rounds, T = the largest surplus or deficit of any bank: extras
// and absences are both dealt round-robin over those T rounds, so every round still holds exactly 32
// values, the first 16 - T rounds are conflict-free and the last T are 2-way (random bank counts give
// T ~ 8: 1.5 wavefronts per LDS.U8 instead of the 3.6 of an unordered unit; tests/test_abi.py).
// perm[] maps an arranged position back to the value's offset inside the unit.
void arrange_fingerprints(const unsigned long long *ts, long long n_vals, long long n_units, unsigned short *fp,
                          unsigned short *perm) {
    std::vector<unsigned short> bank_items[32];
    for (long long u = 0; u < std::max<long long>(1, n_units); ++u) {
        const long long base = u * kFpPerUnit;
        unsigned short f[kFpPerUnit];
        for (int i = 0; i < kFpPerUnit; ++i)
            f[i] = base + i < n_vals ? static_cast<unsigned short>(filter_hash(ts[base + i])) : 0;
        for (auto &b : bank_items) b.clear();
        for (int i = 0; i < kFpPerUnit; ++i) bank_items[(f[i] >> 2) & 31].push_back(static_cast<unsigned short>(i));
        int T = 0;
        for (int b = 0; b < 32; ++b) T = std::max(T, std::abs(static_cast<int>(bank_items[b].size()) - 16));
        T = std::min(T, 16);
        int mult[32][16];  // how many values of bank b go to round k
        for (int b = 0; b < 32; ++b)
            for (int k = 0; k < 16; ++k) mult[b][k] = 1;
        int pe = 0, pa = 0;  // round-robin positions of the extras / of the absences inside the last T rounds
        for (int b = 0; b < 32 && T > 0; ++b) {
            const int c = static_cast<int>(bank_items[b].size());
            for (int e = 0; e < c - 16; ++e) { ++mult[b][16 - T + pe % T]; ++pe; }
            for (int a2 = 0; a2 < 16 - c; ++a2) { --mult[b][16 - T + pa % T]; ++pa; }
        }
        int fill[16];
        unsigned short round_items[16][32];
        size_t next[32];
        for (int k = 0; k < 16; ++k) fill[k] = 0;
        for (int b = 0; b < 32; ++b) next[b] = 0;
        for (int k = 0; k < 16; ++k)
            for (int b = 0; b < 32; ++b)
                for (int m = 0; m < mult[b][k]; ++m) round_items[k][fill[k]++] = bank_items[b][next[b]++];
        for (int r = 0; r < 16; ++r)
            for (int lane = 0; lane < 32; ++lane) {
                const unsigned short i = round_items[r][lane];
                fp[base + lane * 16 + r] = f[i];
                perm[base + lane * 16 + r] = i;
            }
    }
}

int ensure_query_capacity(tvz_match_ws *ws, int qn) {
    if (qn <= ws->q_cap) return TVZ_OK;
    int cap = std::max(256, ws->q_cap);
    while (cap < qn) cap *= 2;
    if (ws->stage_busy) { TVZ_CUDA(cudaEventSynchronize(ws->staged)); ws->stage_busy = false; }
    if (ws->d_qcanon) cudaFree(ws->d_qcanon);
    if (ws->h_stage) cudaFreeHost(ws->h_stage);
    ws->d_qcanon = nullptr;
    ws->h_stage = nullptr;
    ws->q_cap = 0;
    TVZ_CUDA(cudaMalloc(&ws->d_qcanon, sizeof(unsigned long long) * cap));
    // staging holds, per launch chunk, keys+mult, and once the canonical query
    ws->stage_bytes = sizeof(unsigned long long) * cap * 2 + sizeof(int) * cap + 64;
    TVZ_CUDA(cudaHostAlloc(&ws->h_stage, ws->stage_bytes, cudaHostAllocDefault));
    ws->q_cap = cap;
    return TVZ_OK;
}

}  // namespace

extern "C" {

int tvz_catalog_create(const double *h_ts, const int64_t *h_off, const int32_t *h_video_id, int64_t n_rows,
                       tvz_catalog **out) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(out, "null out pointer");
    *out = nullptr;
    TVZ_REQUIRE(n_rows >= 0, "negative n_rows");
    TVZ_REQUIRE(n_rows == 0 || (h_off && h_video_id), "null offsets/video ids");
    TVZ_REQUIRE(n_rows == 0 || h_off[0] == 0, "offsets must start at 0");
    for (int64_t r = 0; r < n_rows; ++r)
        TVZ_REQUIRE(h_off[r + 1] >= h_off[r], "offsets must be non-decreasing (row %lld)", (long long)r);
    const int64_t n_in = n_rows ? h_off[n_rows] : 0;
    TVZ_REQUIRE(n_in == 0 || h_ts, "null timestamps");

    // canonicalise: drop NaN, fold -0.0, drop in-row repeats (first occurrence kept)
    std::vector<unsigned long long> ts;
    ts.reserve(static_cast<size_t>(n_in) + 2);
    std::vector<long long> off(static_cast<size_t>(n_rows) + 1, 0);
    std::unordered_set<unsigned long long> seen;
    for (int64_t r = 0; r < n_rows; ++r) {
        const size_t start = ts.size();
        bool ascending = true;
        double last = 0;
        for (int64_t j = h_off[r]; j < h_off[r + 1]; ++j) {
            const double v = h_ts[j];
            const unsigned long long b = canon_bits(v);
            if (is_nan_bits(b)) continue;
            if (ts.size() > start && !(v > last)) ascending = false;
            last = v;
            ts.push_back(b);
        }
        if (!ascending) {  // rare: unsorted or repeated values -> stable de-duplication
            seen.clear();
            size_t w = start;
            for (size_t j = start; j < ts.size(); ++j)
                if (seen.insert(ts[j]).second) ts[w++] = ts[j];
            ts.resize(w);
        }
        off[r + 1] = static_cast<long long>(ts.size());
    }
    tvz_catalog *c = new tvz_catalog();
    c->n_rows = n_rows;
    c->n_vals = static_cast<long long>(ts.size());
    c->n_pairs = (c->n_vals + 1) / 2;
    c->n_pairs_padded = (c->n_pairs + kChunkPairs - 1) / kChunkPairs * kChunkPairs;
    ts.resize(static_cast<size_t>(2 * c->n_pairs_padded), kPadPattern);
    // coarse index: last row whose offset is <= b*1024 (clamped to the last row)
    std::vector<int> block_row(static_cast<size_t>((2 * c->n_pairs_padded) >> kBlockShift) + 2, 0);
    {
        long long r = 0;
        for (size_t b = 0; b < block_row.size(); ++b) {
            const long long e = static_cast<long long>(b) << kBlockShift;
            while (r + 1 < n_rows && off[r + 1] <= e) ++r;
            block_row[b] = static_cast<int>(r);
        }
    }
    // 16-bit fingerprints, padded to whole warp units (pad entries point past n_vals and are dropped)
    c->n_units = (c->n_vals + kFpPerUnit - 1) / kFpPerUnit;
    std::vector<unsigned short> fp(static_cast<size_t>(std::max<long long>(1, c->n_units)) * kFpPerUnit, 0);
    std::vector<unsigned short> perm(fp.size(), 0);
    arrange_fingerprints(ts.data(), c->n_vals, c->n_units, fp.data(), perm.data());
    // verification records in ARRANGED order: a surviving fingerprint at position p is checked against
    // rec[p].ts and, if it is a real match, adds into counts[rec[p].row] -- one 16-byte load, no search
    // for the row.  Pad positions hold a NaN pattern that equals no query key.
    std::vector<VerifyRec> rec(fp.size(), VerifyRec{kPadPattern, 0u, 0u});
    {
        std::vector<unsigned> row_of(kFpPerUnit);
        long long r = 0;
        for (long long u = 0; u < c->n_units; ++u) {
            const long long base = u * kFpPerUnit;
            for (int i = 0; i < kFpPerUnit && base + i < c->n_vals; ++i) {
                while (r + 1 < n_rows && off[r + 1] <= base + i) ++r;
                row_of[i] = static_cast<unsigned>(r);
            }
            for (int p2 = 0; p2 < kFpPerUnit; ++p2) {
                const long long elem = base + perm[base + p2];
                if (elem < c->n_vals) {
                    rec[base + p2].ts = ts[elem];
                    rec[base + p2].row = row_of[perm[base + p2]];
                }
            }
        }
    }
    cudaGetDevice(&c->device);
    auto fail = [&](cudaError_t e, const char *what) {
        set_error(TVZ_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        tvz_catalog_destroy(c);
        return TVZ_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&c->d_ts, ts.size() * 8)) != cudaSuccess) return fail(e, "cudaMalloc(ts)");
    if ((e = cudaMalloc(&c->d_fp, fp.size() * 2)) != cudaSuccess) return fail(e, "cudaMalloc(fp)");
    if ((e = cudaMemcpy(c->d_fp, fp.data(), fp.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(fp)");
    if ((e = cudaMalloc(&c->d_rec, rec.size() * sizeof(VerifyRec))) != cudaSuccess) return fail(e, "cudaMalloc(rec)");
    if ((e = cudaMemcpy(c->d_rec, rec.data(), rec.size() * sizeof(VerifyRec), cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(rec)");
    if ((e = cudaMalloc(&c->d_off, off.size() * 8)) != cudaSuccess) return fail(e, "cudaMalloc(off)");
    if ((e = cudaMalloc(&c->d_vid, std::max<size_t>(1, n_rows) * 4)) != cudaSuccess) return fail(e, "cudaMalloc(vid)");
    if ((e = cudaMalloc(&c->d_block_row, block_row.size() * 4)) != cudaSuccess) return fail(e, "cudaMalloc(block_row)");
    if ((e = cudaMemcpy(c->d_block_row, block_row.data(), block_row.size() * 4, cudaMemcpyHostToDevice)) !=
        cudaSuccess)
        return fail(e, "cudaMemcpy(block_row)");
    if ((e = cudaMemcpy(c->d_ts, ts.data(), ts.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(ts)");
    if ((e = cudaMemcpy(c->d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(off)");
    if (n_rows && (e = cudaMemcpy(c->d_vid, h_video_id, n_rows * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(vid)");
    *out = c;
    return TVZ_OK;
    });
}

void tvz_catalog_destroy(tvz_catalog *c) {
    if (!c) return;
    if (c->d_ts) cudaFree(c->d_ts);
    if (c->d_fp) cudaFree(c->d_fp);
    if (c->d_rec) cudaFree(c->d_rec);
    if (c->d_off) cudaFree(c->d_off);
    if (c->d_vid) cudaFree(c->d_vid);
    if (c->d_block_row) cudaFree(c->d_block_row);
    delete c;
}

int64_t tvz_catalog_rows(const tvz_catalog *c) { return c ? c->n_rows : 0; }
int64_t tvz_catalog_values(const tvz_catalog *c) { return c ? c->n_vals : 0; }
int64_t tvz_catalog_algo_bytes(const tvz_catalog *c) { return c ? 8 * c->n_vals + 8 * (c->n_rows + 1) : 0; }

int tvz_match_ws_create(const tvz_catalog *cat, int64_t hit_capacity, tvz_match_ws **out) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(cat && out, "null pointer");
    *out = nullptr;
    TVZ_REQUIRE(hit_capacity >= 0, "negative capacity");
    tvz_match_ws *ws = new tvz_match_ws();
    ws->cat = cat;
    ws->cap = std::max<long long>(1, hit_capacity);
    ws->n_blocks = static_cast<int>((cat->n_rows + kScanRowsPerBlock - 1) / kScanRowsPerBlock);
    auto bail = [&](cudaError_t e, const char *what) {
        set_error(TVZ_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        tvz_match_ws_destroy(ws);
        return TVZ_ERR_CUDA;
    };
    cudaError_t e;
    const size_t nr = std::max<long long>(1, cat->n_rows);
    if ((e = cudaMalloc(&ws->d_counts, nr * 4)) != cudaSuccess) return bail(e, "cudaMalloc(counts)");
    if ((e = cudaMemset(ws->d_counts, 0, nr * 4)) != cudaSuccess) return bail(e, "cudaMemset(counts)");
    if ((e = cudaMalloc(&ws->d_state, std::max(1, ws->n_blocks) * 8)) != cudaSuccess) return bail(e, "cudaMalloc(state)");
    if ((e = cudaMemset(ws->d_state, 0, std::max(1, ws->n_blocks) * 8)) != cudaSuccess) return bail(e, "cudaMemset(state)");
    if ((e = cudaMalloc(&ws->d_ticket, 12)) != cudaSuccess) return bail(e, "cudaMalloc(ticket)");
    {
        const unsigned init[3] = {0u, 1u, 0u};
        if ((e = cudaMemcpy(ws->d_ticket, init, 12, cudaMemcpyHostToDevice)) != cudaSuccess)
            return bail(e, "cudaMemcpy(ticket)");
    }
    if ((e = cudaMalloc(&ws->d_out, (ws->cap + 1) * 8)) != cudaSuccess) return bail(e, "cudaMalloc(out)");
    if ((e = cudaMemset(ws->d_out, 0, 8)) != cudaSuccess) return bail(e, "cudaMemset(out)");
    if ((e = cudaMalloc(&ws->d_rows, ws->cap * 8)) != cudaSuccess) return bail(e, "cudaMalloc(rows)");
    if ((e = cudaMalloc(&ws->d_kth, ws->cap * 4)) != cudaSuccess) return bail(e, "cudaMalloc(kth)");
    if ((e = cudaMalloc(&ws->d_nhits, 8)) != cudaSuccess) return bail(e, "cudaMalloc(nhits)");
    if ((e = cudaMalloc(&ws->d_chunk_hits, (nr / kFusedChunk + 2) * 4)) != cudaSuccess)
        return bail(e, "cudaMalloc(chunk_hits)");
    if ((e = cudaMemset(ws->d_nhits, 0, 8)) != cudaSuccess) return bail(e, "cudaMemset(nhits)");
    if ((e = cudaMalloc(&ws->d_keys, kMaxKeys * 8)) != cudaSuccess) return bail(e, "cudaMalloc(keys)");
    if ((e = cudaMalloc(&ws->d_mult, kMaxKeys * 4)) != cudaSuccess) return bail(e, "cudaMalloc(mult)");
    if ((e = cudaHostAlloc(&ws->h_out, (ws->cap + 1) * 8, cudaHostAllocDefault)) != cudaSuccess)
        return bail(e, "cudaHostAlloc(out)");
    if ((e = cudaHostAlloc(&ws->h_kth, ws->cap * 4, cudaHostAllocDefault)) != cudaSuccess)
        return bail(e, "cudaHostAlloc(kth)");
    if ((e = cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&ws->staged, cudaEventDisableTiming)) != cudaSuccess)
        return bail(e, "cudaEventCreate");
    int rc = ensure_query_capacity(ws, 256);
    if (rc) { tvz_match_ws_destroy(ws); return rc; }
    // the memsets above ran on the legacy stream; queries run on non-blocking streams
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bail(e, "cudaDeviceSynchronize");
    *out = ws;
    return TVZ_OK;
    });
}

void tvz_match_ws_destroy(tvz_match_ws *ws) {
    if (!ws) return;
    if (ws->stream) cudaStreamSynchronize(ws->stream);
    if (ws->d_counts) cudaFree(ws->d_counts);
    if (ws->d_state) cudaFree(ws->d_state);
    if (ws->d_ticket) cudaFree(ws->d_ticket);
    if (ws->d_out) cudaFree(ws->d_out);
    if (ws->d_rows) cudaFree(ws->d_rows);
    if (ws->d_kth) cudaFree(ws->d_kth);
    if (ws->d_nhits) cudaFree(ws->d_nhits);
    if (ws->d_chunk_hits) cudaFree(ws->d_chunk_hits);
    if (ws->d_keys) cudaFree(ws->d_keys);
    if (ws->d_mult) cudaFree(ws->d_mult);
    if (ws->d_qcanon) cudaFree(ws->d_qcanon);
    if (ws->h_stage) cudaFreeHost(ws->h_stage);
    if (ws->h_out) cudaFreeHost(ws->h_out);
    if (ws->h_kth) cudaFreeHost(ws->h_kth);
    void *bdev[] = {ws->b_counts, ws->b_out, ws->b_mult, ws->b_nkeys, ws->b_rows, ws->b_nhits, ws->b_state, ws->b_keys,
                    ws->b_ticket};
    for (void *p : bdev)
        if (p) cudaFree(p);
    if (ws->hb_stage) cudaFreeHost(ws->hb_stage);
    if (ws->hb_out) cudaFreeHost(ws->hb_out);
    if (ws->staged) cudaEventDestroy(ws->staged);
    if (ws->t0) cudaEventDestroy(ws->t0);
    if (ws->t1) cudaEventDestroy(ws->t1);
    if (ws->stream) cudaStreamDestroy(ws->stream);
    delete ws;
}

const int32_t *tvz_match_ws_hits(const tvz_match_ws *ws) { return ws ? ws->d_out : nullptr; }
const int64_t *tvz_match_ws_nhits(const tvz_match_ws *ws) {
    return ws ? reinterpret_cast<const int64_t *>(ws->d_nhits) : nullptr;
}
const int32_t *tvz_match_ws_counts(const tvz_match_ws *ws) { return ws ? ws->d_counts : nullptr; }

}  // extern "C"

namespace {

// Enqueue query upload + count + ordered compaction (+ kth) on `st`.  `want_kth` needs qn <= q_cap.
int enqueue_match(const tvz_catalog *cat, tvz_match_ws *ws, const double *h_q, int qn, int min_match, bool want_kth,
                  int *d_out, long long out_cap, cudaStream_t st, const GatherTargets *gather = nullptr) {
    TVZ_REQUIRE(cat && ws && ws->cat == cat, "workspace does not belong to this catalogue");
    TVZ_REQUIRE(qn >= 0 && (qn == 0 || h_q), "bad query");
    if (!d_out) {
        d_out = ws->d_out;
        if (out_cap <= 0) out_cap = ws->cap;
    }
    TVZ_REQUIRE(out_cap >= 1 && out_cap <= ws->cap, "output capacity %lld outside [1, %lld]", out_cap, ws->cap);
    int rc = ensure_query_capacity(ws, std::max(qn, 1));
    if (rc) return rc;
    if (ws->stage_busy) { TVZ_CUDA(cudaEventSynchronize(ws->staged)); ws->stage_busy = false; }

    // canonical query (query order) + sorted distinct keys with multiplicities
    unsigned long long *h_qc = reinterpret_cast<unsigned long long *>(ws->h_stage);
    unsigned long long *h_keys = h_qc + ws->q_cap;
    int *h_mult = reinterpret_cast<int *>(h_keys + ws->q_cap);
    std::vector<unsigned long long> sorted;
    sorted.reserve(qn);
    for (int i = 0; i < qn; ++i) {
        const unsigned long long b = canon_bits(h_q[i]);
        h_qc[i] = b;
        if (!is_nan_bits(b)) sorted.push_back(b);
    }
    std::sort(sorted.begin(), sorted.end());
    int nk = 0;
    for (size_t i = 0; i < sorted.size();) {
        size_t j = i;
        while (j < sorted.size() && sorted[j] == sorted[i]) ++j;
        h_keys[nk] = sorted[i];
        h_mult[nk] = static_cast<int>(j - i);
        ++nk;
        i = j;
    }
    // The pinned staging buffer is read by the device only when a copy out of it is enqueued; the
    // common case (a short query riding in the kernel parameters) never is, so back-to-back
    // asynchronous queries pipeline on the stream without a host-side wait in between.
    bool staged_copy = false;
    if (want_kth && qn > 0) {
        TVZ_CUDA(cudaMemcpyAsync(ws->d_qcanon, h_qc, sizeof(unsigned long long) * qn, cudaMemcpyHostToDevice, st));
        staged_copy = true;
    }
    if (cat->n_rows > 0) {
        const int sms = num_sms();
        const long long want = (cat->n_units + kFpWarps - 1) / kFpWarps;   // at least one unit per warp
        const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(want, static_cast<long long>(TVZ_FP_MINB) * sms)));
        if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t0, st));
        // One launch for the whole query when the keys fit one count launch: the kernel compacts its own
        // counts behind two grid-wide barriers (cooperative launch).  TVZ_NO_FUSE=1 keeps the two kernels.
        static const bool fuse_ok = [] {
            const char *e = getenv("TVZ_NO_FUSE");
            return !(e && e[0] == '1');
        }();
        bool fused = fuse_ok && nk > 0 && nk <= kMaxKeys;
        FusedCompact fc;
        if (fused) {
            fc.enabled = 1;
            fc.min_match = min_match;
            fc.n_rows = cat->n_rows;
            fc.cap = out_cap;
            fc.vid = cat->d_vid;
            fc.out = d_out;
            fc.rows_out = ws->d_rows;
            fc.n_hits_out = ws->d_nhits;
            fc.chunk_hits = ws->d_chunk_hits;
            fc.done = ws->d_ticket + 2;
            if (gather) fc.gt = *gather;
        }
        auto launch = [&](bool param, const unsigned long long *dk, const int *dm, int n, const SmallQuery &sq) -> int {
            auto kern = param ? match_count_kernel<true> : match_count_kernel<false>;
            TVZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(sizeof(FpSmem))));
            cudaError_t le = launch_pdl(kern, dim3(fused ? 2 * sms : grid), dim3(kFpThreads), sizeof(FpSmem), st, fused,
                                        cat->d_fp, cat->n_units, cat->d_rec, dk, dm, n, ws->d_counts, sq, fc);
            if (le == cudaErrorCooperativeLaunchTooLarge && fused) {
                // the device cannot hold 2 CTAs per SM right now (MPS limits, a debugger ...): two kernels
                cudaGetLastError();
                fused = false;
                fc.enabled = 0;
                le = launch_pdl(kern, dim3(grid), dim3(kFpThreads), sizeof(FpSmem), st, false, cat->d_fp, cat->n_units,
                                cat->d_rec, dk, dm, n, ws->d_counts, sq, fc);
            }
            TVZ_CUDA(le);
            return TVZ_OK;
        };
        if (nk <= kParamKeys) {
            SmallQuery sq;
            memcpy(sq.keys, h_keys, sizeof(unsigned long long) * nk);
            memcpy(sq.mult, h_mult, sizeof(int) * nk);
            if (nk > 0) {
                rc = launch(true, nullptr, nullptr, nk, sq);
                if (rc) return rc;
            }
        } else {
            for (int k0 = 0; k0 < nk; k0 += kMaxKeys) {
                const int n = std::min(kMaxKeys, nk - k0);
                TVZ_CUDA(cudaMemcpyAsync(ws->d_keys, h_keys + k0, sizeof(unsigned long long) * n,
                                         cudaMemcpyHostToDevice, st));
                TVZ_CUDA(cudaMemcpyAsync(ws->d_mult, h_mult + k0, sizeof(int) * n, cudaMemcpyHostToDevice, st));
                staged_copy = true;
                rc = launch(false, ws->d_keys, ws->d_mult, n, SmallQuery{});
                if (rc) return rc;
            }
        }
        if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t1, st));
        if (!fused) {
            rc = compact_enqueue(ws->d_counts, cat->n_rows, min_match, cat->d_vid, d_out, ws->d_rows, out_cap,
                                 ws->d_nhits, ws->d_state, ws->d_ticket, nullptr, nullptr, st, gather);
            if (rc) return rc;
        }
        if (want_kth) {
            match_kth_kernel<<<2 * sms, 256, 0, st>>>(cat->d_ts, cat->d_off, ws->d_rows, d_out, out_cap,
                                                      ws->d_qcanon, qn, min_match, ws->d_kth);
            TVZ_CUDA(cudaGetLastError());
        }
    } else {
        TVZ_REQUIRE(!gather || gather->n_peers == 0, "an empty shard cannot take part in the fused gather");
        TVZ_CUDA(cudaMemsetAsync(d_out, 0, 8, st));
        TVZ_CUDA(cudaMemsetAsync(ws->d_nhits, 0, 8, st));
    }
    if (staged_copy) {
        TVZ_CUDA(cudaEventRecord(ws->staged, st));
        ws->stage_busy = true;
    }
    return TVZ_OK;
}

}  // namespace

extern "C" {

// Debug hooks (not in the public header): time the count kernel of the last query with CUDA
// events recorded on the query's own stream.
int tvz_debug_match_timing(tvz_match_ws *ws, int enable) {
    TVZ_REQUIRE(ws, "null workspace");
    if (enable && !ws->t0) {
        TVZ_CUDA(cudaEventCreate(&ws->t0));
        TVZ_CUDA(cudaEventCreate(&ws->t1));
    }
    ws->timing = enable != 0;
    return TVZ_OK;
}
int tvz_debug_match_count_ms(tvz_match_ws *ws, float *ms) {
    TVZ_REQUIRE(ws && ms && ws->t0, "timing was never enabled");
    TVZ_CUDA(cudaEventSynchronize(ws->t1));
    TVZ_CUDA(cudaEventElapsedTime(ms, ws->t0, ws->t1));
    return TVZ_OK;
}

// Debug hook (host only, no GPU needed): the fingerprint layout of `n` stored values, as the catalogue
// packer computes it.  fp_out / perm_out hold ceil(n / 512) * 512 entries.
int tvz_debug_arrange_fingerprints(const double *values, int64_t n, uint16_t *fp_out, uint16_t *perm_out) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(n >= 0 && (n == 0 || values) && fp_out && perm_out, "bad arguments");
    std::vector<unsigned long long> bits(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) bits[i] = canon_bits(values[i]);
    arrange_fingerprints(bits.data(), n, (n + kFpPerUnit - 1) / kFpPerUnit, fp_out, perm_out);
    return TVZ_OK;
    });
}

int tvz_catalog_match_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *h_q, int qn, int min_match,
                            int32_t *d_out, int64_t out_cap, void *stream) {
    return guarded([&]() -> int {
    return enqueue_match(cat, ws, h_q, qn, min_match, false, d_out, out_cap, static_cast<cudaStream_t>(stream));
    });
}

int tvz_catalog_match_gather_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *h_q, int qn,
                                   int min_match, int n_peers, const uint64_t *peer_record, const uint64_t *peer_flag,
                                   const uint32_t *d_my_flags, int64_t out_cap, uint32_t epoch, void *stream) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers, "n_peers %d outside [1, %d]", n_peers, kMaxPeers);
    TVZ_REQUIRE(peer_record && peer_flag && d_my_flags, "null pointer");
    TVZ_REQUIRE(cat && cat->n_rows > 0, "the fused gather needs a non-empty shard");
    GatherTargets gt;
    gt.n_peers = n_peers;
    gt.epoch = epoch;
    for (int p = 0; p < n_peers; ++p) {
        gt.record[p] = reinterpret_cast<int *>(static_cast<uintptr_t>(peer_record[p]));
        gt.flag[p] = reinterpret_cast<unsigned *>(static_cast<uintptr_t>(peer_flag[p]));
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = enqueue_match(cat, ws, h_q, qn, min_match, false, nullptr, out_cap, st, &gt);
    if (rc) return rc;
    return gather_wait_enqueue(d_my_flags, n_peers, epoch, st);
    });
}

int tvz_catalog_match(const tvz_catalog *cat, tvz_match_ws *ws, const double *q, int qn, int min_match,
                      int32_t *out_video_id, int32_t *out_count, int32_t *out_kth, int64_t cap, int64_t *n_out) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(n_out, "null n_out");
    *n_out = 0;
    TVZ_REQUIRE(cap >= 0 && (cap == 0 || (out_video_id && out_count)), "bad output buffers");
    int rc = enqueue_match(cat, ws, q, qn, min_match, out_kth != nullptr, nullptr, 0, ws ? ws->stream : nullptr);
    if (rc) return rc;
    cudaStream_t st = ws->stream;
    // header + an optimistic first slice of the hit list in one copy
    const long long first = std::min<long long>(ws->cap, 2048);
    TVZ_CUDA(cudaMemcpyAsync(ws->h_out, ws->d_out, (first + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (out_kth) TVZ_CUDA(cudaMemcpyAsync(ws->h_kth, ws->d_kth, first * 4, cudaMemcpyDeviceToHost, st));
    TVZ_CUDA(cudaStreamSynchronize(st));
    ws->stage_busy = false;
    long long n_hits = ws->h_out[0];
    if (n_hits == 0x7fffffff) {
        TVZ_CUDA(cudaMemcpy(&n_hits, ws->d_nhits, 8, cudaMemcpyDeviceToHost));
    }
    *n_out = n_hits;
    if (n_hits > ws->cap || n_hits > cap)
        return set_error(TVZ_ERR_OVERFLOW, "%lld rows qualify; workspace capacity %lld, caller capacity %lld",
                         n_hits, ws->cap, (long long)cap);
    if (n_hits > first) {
        TVZ_CUDA(cudaMemcpyAsync(ws->h_out + 2 * (first + 1), ws->d_out + 2 * (first + 1), (n_hits - first) * 8,
                                 cudaMemcpyDeviceToHost, st));
        if (out_kth)
            TVZ_CUDA(cudaMemcpyAsync(ws->h_kth + first, ws->d_kth + first, (n_hits - first) * 4,
                                     cudaMemcpyDeviceToHost, st));
        TVZ_CUDA(cudaStreamSynchronize(st));
    }
    for (long long h = 0; h < n_hits; ++h) {
        out_video_id[h] = ws->h_out[2 + 2 * h];
        out_count[h] = ws->h_out[3 + 2 * h];
    }
    if (out_kth) memcpy(out_kth, ws->h_kth, n_hits * 4);
    return TVZ_OK;
    });
}

}  // extern "C"

namespace {

int ensure_batch_buffers(tvz_match_ws *ws) {
    const tvz_catalog *cat = ws->cat;
    if (ws->b_counts && ws->hb_cap == ws->cap) return TVZ_OK;
    TVZ_REQUIRE(!ws->b_counts, "batch buffers cannot be resized");  // cap is fixed per workspace
    const size_t nr = std::max<long long>(1, cat->n_rows);
    const size_t nb = std::max(1, ws->n_blocks);
    TVZ_CUDA(cudaMalloc(&ws->b_counts, kBatch * nr * 4));
    TVZ_CUDA(cudaMemset(ws->b_counts, 0, kBatch * nr * 4));
    TVZ_CUDA(cudaMalloc(&ws->b_out, kBatch * (ws->cap + 1) * 8));
    TVZ_CUDA(cudaMalloc(&ws->b_rows, kBatch * ws->cap * 8));
    TVZ_CUDA(cudaMalloc(&ws->b_nhits, kBatch * 8));
    TVZ_CUDA(cudaMalloc(&ws->b_state, kBatch * nb * 8));
    TVZ_CUDA(cudaMemset(ws->b_state, 0, kBatch * nb * 8));
    TVZ_CUDA(cudaMalloc(&ws->b_ticket, kBatch * 16));
    unsigned init[kBatch * 4];
    for (int b = 0; b < kBatch; ++b) { init[4 * b] = 0; init[4 * b + 1] = 1; init[4 * b + 2] = 0; init[4 * b + 3] = 0; }
    TVZ_CUDA(cudaMemcpy(ws->b_ticket, init, sizeof init, cudaMemcpyHostToDevice));
    TVZ_CUDA(cudaMalloc(&ws->b_keys, kBatch * kParamKeys * 8));
    TVZ_CUDA(cudaMalloc(&ws->b_mult, kBatch * kParamKeys * 4));
    TVZ_CUDA(cudaMalloc(&ws->b_nkeys, kBatch * 4));
    TVZ_CUDA(cudaHostAlloc(&ws->hb_stage, kBatch * kParamKeys * 12 + kBatch * 4, cudaHostAllocDefault));
    TVZ_CUDA(cudaHostAlloc(&ws->hb_out, kBatch * (ws->cap + 1) * 8, cudaHostAllocDefault));
    TVZ_CUDA(cudaDeviceSynchronize());
    ws->hb_cap = ws->cap;
    return TVZ_OK;
}

}  // namespace

extern "C" {

int tvz_catalog_batch_limit(void) { return kParamKeys; }

/* Up to 8 queries per catalogue pass; more are processed group by group.  Queries with more than
 * tvz_catalog_batch_limit() distinct values are refused (run them through tvz_catalog_match). */
int tvz_catalog_match_batch(const tvz_catalog *cat, tvz_match_ws *ws, const double *q_all, const int64_t *q_off,
                            int n_queries, int min_match, int32_t *out_video_id, int32_t *out_count,
                            int64_t *out_off, int64_t cap_total, int64_t *need_per_query, int64_t *need_total_out) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(cat && ws && ws->cat == cat, "workspace does not belong to this catalogue");
    TVZ_REQUIRE(n_queries >= 0 && q_off && out_off && need_per_query && need_total_out, "bad arguments");
    *need_per_query = 0;
    *need_total_out = 0;
    TVZ_REQUIRE(cap_total >= 0 && (cap_total == 0 || (out_video_id && out_count)), "bad output buffers");
    out_off[0] = 0;
    if (n_queries == 0) return TVZ_OK;
    int rc = ensure_batch_buffers(ws);
    if (rc) return rc;
    cudaStream_t st = ws->stream;
    unsigned long long *h_keys = reinterpret_cast<unsigned long long *>(ws->hb_stage);
    int *h_mult = reinterpret_cast<int *>(h_keys + kBatch * kParamKeys);
    int *h_nk = h_mult + kBatch * kParamKeys;
    const long long rec = (ws->cap + 1) * 2;
    long long written = 0;
    bool overflow = false;
    long long need_cap = 0, need_total = 0;
    std::vector<unsigned long long> sorted;
    for (int g0 = 0; g0 < n_queries; g0 += kBatch) {
        const int nb = std::min(kBatch, n_queries - g0);
        for (int b = 0; b < nb; ++b) {
            const double *q = q_all + q_off[g0 + b];
            const long long qn = q_off[g0 + b + 1] - q_off[g0 + b];
            TVZ_REQUIRE(qn >= 0, "query offsets must be non-decreasing");
            sorted.clear();
            for (long long i = 0; i < qn; ++i) {
                const unsigned long long bits = canon_bits(q[i]);
                if (!is_nan_bits(bits)) sorted.push_back(bits);
            }
            std::sort(sorted.begin(), sorted.end());
            int nk = 0;
            for (size_t i = 0; i < sorted.size();) {
                size_t j = i;
                while (j < sorted.size() && sorted[j] == sorted[i]) ++j;
                TVZ_REQUIRE(nk < kParamKeys, "query %d has more than %d distinct values: not batchable", g0 + b,
                            kParamKeys);
                h_keys[b * kParamKeys + nk] = sorted[i];
                h_mult[b * kParamKeys + nk] = static_cast<int>(j - i);
                ++nk;
                i = j;
            }
            h_nk[b] = nk;
        }
        if (cat->n_rows > 0) {
            TVZ_CUDA(cudaMemcpyAsync(ws->b_keys, h_keys, kBatch * kParamKeys * 8, cudaMemcpyHostToDevice, st));
            TVZ_CUDA(cudaMemcpyAsync(ws->b_mult, h_mult, kBatch * kParamKeys * 4, cudaMemcpyHostToDevice, st));
            TVZ_CUDA(cudaMemcpyAsync(ws->b_nkeys, h_nk, kBatch * 4, cudaMemcpyHostToDevice, st));
            TVZ_CUDA(cudaFuncSetAttribute(match_count_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(sizeof(BatchSmem))));
            const long long chunks = cat->n_pairs_padded / kChunkPairs;
            const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(chunks, 2ll * num_sms())));
            if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t0, st));
            match_count_batch_kernel<<<grid, kCountThreads, sizeof(BatchSmem), st>>>(
                reinterpret_cast<const ulonglong2 *>(cat->d_ts), cat->n_pairs_padded, ws->b_keys, ws->b_mult, ws->b_nkeys,
                nb, cat->d_off, cat->d_block_row, cat->n_rows, ws->b_counts, cat->n_rows);
            TVZ_CUDA(cudaGetLastError());
            if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t1, st));
            BatchStrides bs;
            bs.counts = cat->n_rows;
            bs.out = rec;
            bs.rows = ws->cap;
            bs.state = std::max(1, ws->n_blocks);
            rc = compact_enqueue_batch(ws->b_counts, cat->n_rows, min_match, cat->d_vid, ws->b_out, ws->b_rows, ws->cap,
                                       ws->b_nhits, ws->b_state, ws->b_ticket, nb, bs, st);
            if (rc) return rc;
            // headers of the group in one strided copy, then each query's hits
            TVZ_CUDA(cudaMemcpy2DAsync(ws->hb_out, rec * 4, ws->b_out, rec * 4, 8, nb, cudaMemcpyDeviceToHost, st));
            TVZ_CUDA(cudaStreamSynchronize(st));
            for (int b = 0; b < nb; ++b) {
                const long long n = ws->hb_out[b * rec];
                if (n > ws->cap) { overflow = true; need_cap = std::max(need_cap, n); continue; }
                if (n > 0)
                    TVZ_CUDA(cudaMemcpyAsync(ws->hb_out + b * rec + 2, ws->b_out + b * rec + 2, n * 8,
                                             cudaMemcpyDeviceToHost, st));
            }
            TVZ_CUDA(cudaStreamSynchronize(st));
        } else {
            for (int b = 0; b < nb; ++b) ws->hb_out[b * rec] = 0;
        }
        for (int b = 0; b < nb; ++b) {
            const long long n = std::min<long long>(ws->hb_out[b * rec], ws->cap);
            need_total += ws->hb_out[b * rec];
            if (!overflow && written + n <= cap_total) {
                for (long long h = 0; h < n; ++h) {
                    out_video_id[written + h] = ws->hb_out[b * rec + 2 + 2 * h];
                    out_count[written + h] = ws->hb_out[b * rec + 3 + 2 * h];
                }
                written += n;
            } else {
                overflow = true;
            }
            out_off[g0 + b + 1] = written;
        }
    }
    *need_per_query = need_cap;
    *need_total_out = need_total;
    if (overflow) {
        return set_error(TVZ_ERR_OVERFLOW, "batch results need %lld entries in total (caller gave %lld) and %lld per "
                         "query (workspace holds %lld)", need_total, (long long)cap_total, need_cap, ws->cap);
    }
    return TVZ_OK;
    });
}

}  // extern "C"
