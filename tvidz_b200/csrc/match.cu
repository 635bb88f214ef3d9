// Stage 2: find_duplicates over a device-resident packed catalogue.
//
// Replaces inspector/db.py:76-94 (full-table fetch at db.py:83 + the O(N*q*L) Python
// membership loop at db.py:85-91) and the row upsert of db.py:43-64.  Semantics:
//     match_count(row) = #{ i : q[i] == some element of row }      (float ==)
//     result = [(video_id, match_count) for rows with match_count >= min_match]
//
// Data layout in HBM (one shard per GPU):
//     fp   : uint16 [units * 512]  filter_hash of every stored value -- what a query streams (2 B per
//            value); inside each 512-value unit the order is chosen for conflict-free lookups
//     rec  : {u64 value, u32 row} [units * 512] in the same order -- what a surviving fingerprint is
//            verified against (IEEE-754 bit patterns, canonicalised at pack time: -0.0 -> +0.0, NaN
//            dropped, in-row repeats dropped, so bitwise equality == Python float equality and every
//            stored value adds at most once)
//     ts   : u64 [values] in row order + off i64 [rows + 1] (CSR): the per-cut early-exit kernel
//     vid  : int32 [rows]; dead : u8 [rows] (rows replaced by an upsert)
//     tiles: {row_lo, n_rows, unit_lo, unit_hi}: the catalogue cut into <= 4096-row pieces of equal
//            stored-value count, one CTA each
// Because in-row repeats are gone, match_count(row) = sum over stored values v of mult(v), where
// mult(v) = number of query positions equal to v.
//
// ONE kernel per query (match_tile_kernel): a CTA owns WHOLE ROWS, so their counts live in its shared
// memory -- no global counts[] array to zero, re-read and compact, no grid-wide barrier, no
// cooperative launch.  It streams the fingerprints of its rows (256-bit loads, a 64 KB byte map of
// the query rejects ~all values with one LDS.U8), verifies the rare survivors against `rec`, then
// compacts its own rows: qualifying rows are counted, the tile's total is published, the totals of
// all earlier tiles are read in ONE parallel step (every thread polls one predecessor: a single
// global round trip instead of a look-back chain), and the hits go to their final position in
// catalogue order.  With the fused multi-GPU gather the same kernel stores every hit into all peers'
// buffers over NVLink and its last CTA publishes header + flag and waits for the peers' flags.
// The kernel is templated on the number of queries answered per pass (1, or up to 8: one bit per
// query in the byte map, 16-bit counts).
//
// The last tile is the mutable TAIL: rows written by tvz_catalog_upsert (one small kernel, keys in its
// parameters) go there, the row they replace is neutralised in place (its verification records are
// overwritten with a NaN pattern that equals no query value), so the reference's per-cut
// add_timestamps() + find_duplicates() loop (app.py:234-235) never repacks the catalogue.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace tvz {
namespace {

constexpr int kMapEntries = 1 << 16;         // byte-map filter of the query: 64 KB of shared memory
constexpr int kFpPerUnit = 32 * 16;          // fingerprints per warp-wide 256-bit load
constexpr int kTileRows = 4096;              // rows whose counts one CTA keeps in shared memory
constexpr int kTileKeys = 1024;              // distinct values of ONE query kept in shared memory (more: searched in global memory)
constexpr int kParamKeys = 224;              // distinct values that ride in the kernel parameters; per-query limit of a batch
constexpr int kBatch = 8;                    // queries per batched pass
constexpr int kRecvCtas = 32;                // fused gather: CTAs (the grid's last) that wait for the peers' records
constexpr int kMinTileUnits = 32;            // small catalogues: at least two units per warp and tile
constexpr int kFpPadUnits = 64;              // fingerprint array padding: a tile's first (speculative) loads stay in bounds
constexpr int kMaxTailTiles = 16;            // the mutable tail: up to 16 tiles = 65536 rows between repacks
constexpr int kTailRows = kMaxTailTiles * kTileRows;
constexpr unsigned long long kPadPattern = 0x7ff8dead0000beefull;  // a NaN: never equals a stored value

// Two IMADs and a shift: good enough on frame-quantised timestamps and on x.0 / x.5 values
// (false-positive rate ~ n_keys / 65536, measured in DESIGN.md).
__host__ __device__ __forceinline__ uint32_t filter_hash(unsigned long long v) {
    const uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
    return (lo * 0x9E3779B1u + hi * 0x85EBCA77u) >> 16;
}

// Short queries (the common case: a video has tens of cuts) ride in the kernel parameters, which saves the
// host->device copy in front of the launch -- and with it the copy engine between two kernels of a stream, which
// would also end the launch overlap (PDL) of consecutive queries.  kP = distinct values per query the parameter block
// holds; 0 = the keys are in global memory (a long single query).  A batch of 8 comes in two sizes: kBatchShortKeys for
// ordinary cut lists (9 KB of parameters), kParamKeys (21 KB; the limit is 32,764 bytes) for everything batchable.
constexpr int kBatchShortKeys = 96;
template <int kQ, int kP>
struct QueryParam {
    unsigned long long keys[kQ][kP];
    int mult[kQ][kP];
    int n_keys[kQ];
};
template <int kQ>
struct QueryParam<kQ, 0> {
    int unused;
};

// 256-bit streaming load (sm_100 LDG.E.256): no L1 allocation -- shared memory takes ~205 of the SM's
// 228 KB -- and evict-first in L2.
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

// One CTA's share of the catalogue: whole rows [row_lo, row_lo + n_rows) and the fingerprint units that
// hold their values.  Neighbouring tiles may share a boundary unit; each counts only its own rows.
struct alignas(16) TileDesc {
    int row_lo, n_rows;
    unsigned unit_lo, unit_hi;
};

// TVZ_Q1_WIDE = 1: the single-query kernel runs ONE 1024-thread CTA per SM on a PAIR of adjacent tiles (half the
// CTAs in the tile-total exchange, one byte map per SM to zero instead of two); 0: one 512-thread CTA per tile, 2 per SM.
#ifndef TVZ_Q1_WIDE
#define TVZ_Q1_WIDE 1   // measured (r02): 1M rows 34.1 vs 37.1 us back to back, 44.6 vs 46.6 us cold; 125k-row shard 20.5 vs 25.1 us cold
#endif
// TVZ_BATCH_WIDE = 1: the 8-query kernel works on tile pairs as well (148 CTAs = ONE wave instead of two: its 128 KB of
// 16-bit counts + the 64 KB byte map leave room for one CTA per SM either way); 0: one tile per CTA.
#ifndef TVZ_BATCH_WIDE
#define TVZ_BATCH_WIDE 1
#endif
// TVZ_BATCH_PARAMS = 1: a batch's keys ride in the kernel parameters; 0: staged through pinned memory + one copy.
#ifndef TVZ_BATCH_PARAMS
#define TVZ_BATCH_PARAMS 1
#endif
// TVZ_STAGED_EMIT: whose hits are ordered in shared memory first and written out by consecutive threads (coalesced
// record stores, contiguous NVLink stores in the fused gather) instead of every thread storing its own.
// 1 = the 8-query pass only (shipped), 2 = single queries too, 0 = nobody.  Measured: 8 queries at 1 M rows 92.0 ->
// 89.4 us per pass, at N = 2 with the gather 109.2 -> 104.8 us; one query at N = 1 is 0.8 us SLOWER staged (two more
// barriers for 19 k hits spread over 148 CTAs), at N = 2 the same within noise (profiles/r02_gather_variants_n2.txt).
#ifndef TVZ_STAGED_EMIT
#define TVZ_STAGED_EMIT 1
#endif
template <int kQ>
struct TileShape {
    static constexpr int kPair = ((kQ == 1 && TVZ_Q1_WIDE) || (kQ > 1 && TVZ_BATCH_WIDE)) ? 2 : 1;   // adjacent tiles one CTA works on
    static constexpr int kThreads = (kQ == 1 && !TVZ_Q1_WIDE) ? 512 : 1024;
    static constexpr int kMinBlocks = (kQ == 1 && !TVZ_Q1_WIDE) ? 2 : 1;
    static constexpr int kWarps = kThreads / 32;
    static constexpr int kQueue = (kQ == 1 && !TVZ_Q1_WIDE) ? 128 : 64;   // survivors parked per warp
    static constexpr int kKeys = kQ == 1 ? kTileKeys : kParamKeys;
    static constexpr int kRows = kTileRows * kPair;               // rows whose counts the CTA keeps
    static constexpr int kCountWords = kQ == 1 ? kRows : kQ * kRows / 2;  // batch: two 16-bit counts per word
    static constexpr int kRowsPerThread = kRows / kThreads;
    static constexpr int kGroup = kThreads / kQ;                  // threads that poll the predecessors of one query
};

template <int kQ>
struct alignas(16) TileSmem {
    using S = TileShape<kQ>;
    unsigned char map[kMapEntries];        // first: zeroed with 16-byte stores
    unsigned long long keys[kQ][S::kKeys];
    unsigned long long part[S::kWarps];    // per-warp partial sums of the predecessors' totals
    unsigned long long excl[kQ];           // hits of all earlier tiles
    alignas(16) unsigned counts[S::kCountWords];   // zeroed with 16-byte stores
    unsigned qe[S::kWarps][S::kQueue];     // parked survivors: arranged position relative to the tile's first unit
    int mult[kQ][S::kKeys];
    int n_keys[kQ];
    unsigned warp_tot[kQ][S::kWarps];      // qualifying rows per warp, then their exclusive prefix
    unsigned agg[kQ];
};

struct TileArgs {
    const unsigned short *fp;
    const VerifyRec *rec;
    const TileDesc *tiles;
    TileDesc tail[kMaxTailTiles];           // the mutable tail's tiles, as they were when the query was enqueued
    int n_tiles, tail_index;                // tiles >= tail_index belong to the tail; -1: none
    const unsigned long long *keys_g;       // keys in global memory: a long single query, or the batch [n_queries][key_stride]
    const int *mult_g;
    const int *n_keys_g;                    // batch: distinct values per query
    int key_stride, n_queries;
    int n_keys;                             // single query
    int min_match;
    long long cap;
    const int *vid;
    const unsigned char *dead;
    int *out;                               // [n_queries][out_stride]: {n_hits, overflow}, then (video_id, match_count)
    long long out_stride;
    long long *rows_out;                    // [cap] row index of every hit (single query; nullable)
    long long *n_hits_out;                  // [n_queries]
    unsigned *state;                        // [n_tiles][kQ]: {query sequence << 16 | qualifying rows of the tile}
    unsigned seq;                           // this query's sequence number on its workspace (1..65535)
    int uniform_units;                      // > 0: packed tile t starts at unit t * uniform_units (no lookup before the first loads)
    long long *trace;                       // debug: [n_tiles][16] phase timestamps (tvz_debug_tile_trace), normally null
    long long n_pos, n_rows_cap;            // checked build: arranged positions / rows the arrays hold
    GatherTargets gt;
};

// L2-coherent accesses to the words CTAs exchange (no L1, no fence of their own)
// Checked build (-DTVZ_CHECKED, scripts/checked_build.sh): every index the tile and upsert kernels compute is
// tested against its bound and a violation traps with a message -- this pool's GPUs refuse compute-sanitizer
// (profiles/r02_sanitizer_refused.txt), so the bounds checks are the library's own.
#ifdef TVZ_CHECKED
#define TVZ_CHECK(cond)                                                                                     \
    do {                                                                                                    \
        if (!(cond)) {                                                                                      \
            printf("tvidz_b200 CHECK failed: %s (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__,  \
                   (int)blockIdx.x, (int)threadIdx.x);                                                      \
            __trap();                                                                                       \
        }                                                                                                   \
    } while (0)
#else
#define TVZ_CHECK(cond) ((void)0)
#endif

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Tagged 16-byte record entries of the fused gather: {v.x, epoch, v.y, epoch}.  The store may be split in
// transit, but each 8-byte half is atomic and carries its own tag.
__device__ __forceinline__ void st_tagged(int *entry, int2 v, unsigned epoch) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(entry), "r"(v.x), "r"(epoch), "r"(v.y) : "memory");
}
// One look at an entry: are both halves there?
__device__ __forceinline__ bool try_tagged(const int *entry, unsigned epoch) {
    [[maybe_unused]] unsigned a0, a1;   // the values: only the tags are looked at here
    unsigned t0, t1;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(t0), "=r"(a1), "=r"(t1) : "l"(entry) : "memory");
    return t0 == epoch && t1 == epoch;
}
// Poll (bounded) until both halves of an entry carry `epoch`; returns the two values.
__device__ __forceinline__ int2 ld_tagged(const int *entry, unsigned epoch) {
    unsigned polls = 0;
    while (true) {
        unsigned a0, t0, a1, t1;
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(a0), "=r"(t0) : "l"(entry) : "memory");
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(a1), "=r"(t1) : "l"(entry + 2) : "memory");
        if (t0 == epoch && t1 == epoch) return make_int2(static_cast<int>(a0), static_cast<int>(a1));
        if (++polls > 1024) __nanosleep(64);
        if (polls == (1u << 25)) __trap();   // a peer that never answers turns into a launch failure, not a hung GPU
    }
}

// How a predecessor's total is polled.  TVZ_POLL_LD: 0 = ld.relaxed.gpu (a strong load per lane), 1 = ld.global.cg
// (L2-only weak load, volatile asm so that the loop re-reads; the warp's lanes coalesce into line requests)
#ifndef TVZ_POLL_LD
#define TVZ_POLL_LD 0   // measured: no difference (125k-row shard, cold 23.4 vs 24.5 us, back to back 14.7 both)
#endif
__device__ __forceinline__ unsigned ld_poll_u32(const unsigned *p) {
#if TVZ_POLL_LD == 0
    return ld_relaxed_u32(p);
#else
    unsigned v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Start the fetch of a survivor's verification record without waiting for it.  TVZ_WARM: 1 = prefetch.global.L2
// (brings the whole 128-byte line, or more), 2 = a 4-byte load whose result is never used (brings the sector),
// 0 = nothing (the verification pays the full latency).
#ifndef TVZ_WARM
#define TVZ_WARM 1   // measured (profiles/r02_warm_variants.txt): same DRAM bytes for all three, 1 and 2 are ~2 us faster than 0
#endif
__device__ __forceinline__ void warm_record(const void *p) {
#if TVZ_WARM == 1
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#elif TVZ_WARM == 2
    unsigned unused;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(unused) : "l"(p));
#else
    (void)p;
#endif
}

template <class T>
__device__ __forceinline__ int lower_bound_u64(const T &at, int n, unsigned long long v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (at(mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int kQ, int kP>
__global__ void __launch_bounds__(TileShape<kQ>::kThreads, TileShape<kQ>::kMinBlocks)
match_tile_kernel(const __grid_constant__ TileArgs a, const __grid_constant__ QueryParam<kQ, kP> sq) {
    using S = TileShape<kQ>;
    constexpr bool kParamQuery = kP > 0;
    static_assert(kP <= S::kKeys, "the parameter block cannot hold more keys than shared memory does");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem<kQ> &sm = *reinterpret_cast<TileSmem<kQ> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;                    // this CTA's place in the exchange (= its tile, or its pair of tiles)
    const int t0 = tile * S::kPair;                 // first tile of the CTA
    const bool last_cta = tile == static_cast<int>(gridDim.x) - 1;
    const int nq = kQ == 1 ? 1 : a.n_queries;
    auto mark = [&](int k) {   // debug trace: slot 0 = global ns at CTA start, slots 1..7 = SM cycles at the phase ends
        if (a.trace && tid == 0) {
            long long t;
            if (k == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            else t = clock64();
            a.trace[static_cast<size_t>(tile) * 16 + k] = t;
        }
    };
    mark(0);
    mark(1);

    // ---- prologue: nothing here depends on the previous kernel of the stream ----
    // The tile descriptor is on its way while the byte map and the counts are zeroed; the first
    // fingerprint loads follow it (they do not depend on the query either).
    const bool in_tail = a.tail_index >= 0 && t0 >= a.tail_index;
    const bool uniform = a.uniform_units > 0 && !in_tail;
    TileDesc td;
    if (in_tail) td = a.tail[t0 - a.tail_index];
    else td = a.tiles[t0];
    if (S::kPair == 2 && t0 + 1 < a.n_tiles) {      // rows and units of adjacent tiles are contiguous: one merged range
        const int t1 = t0 + 1;
        const TileDesc d1 = (a.tail_index >= 0 && t1 >= a.tail_index) ? a.tail[t1 - a.tail_index] : a.tiles[t1];
        td.n_rows += d1.n_rows;
        td.unit_hi = max(td.unit_hi, d1.unit_hi);
    }
    const bool resident = kQ > 1 || a.n_keys <= S::kKeys;   // the query's keys fit shared memory
    const bool scan = kQ > 1 || a.n_keys > 0;               // an empty query matches nothing: no stream
    // Packed tiles start at a unit that is pure arithmetic (tile * uniform_units), so the first
    // fingerprint loads -- which do not depend on the query either -- leave before anything has been
    // read; the array is padded so that they are in bounds even where a tile turns out to be shorter.
    const unsigned unit_lo = uniform ? static_cast<unsigned>(t0) * static_cast<unsigned>(a.uniform_units) : td.unit_lo;
    const unsigned short *lane_fp = a.fp + lane * 16;
    U32x8 v[2];
    if (uniform && scan) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
            v[j] = ld_stream_u32x8(lane_fp + static_cast<size_t>(unit_lo + warp + j * S::kWarps) * kFpPerUnit);
    }
    for (int i = tid; i < kMapEntries / 16; i += S::kThreads)
        reinterpret_cast<uint4 *>(sm.map)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < S::kCountWords / 4; i += S::kThreads)
        reinterpret_cast<uint4 *>(sm.counts)[i] = make_uint4(0u, 0u, 0u, 0u);
    const unsigned unit_hi = scan ? td.unit_hi : unit_lo;
    TVZ_CHECK(td.n_rows >= 0 && td.n_rows <= S::kRows && td.row_lo >= 0 && td.row_lo + td.n_rows <= a.n_rows_cap);
    TVZ_CHECK(td.unit_hi >= td.unit_lo && static_cast<long long>(td.unit_hi) * kFpPerUnit <= a.n_pos);
    TVZ_CHECK(!uniform || td.unit_lo == unit_lo);
    TVZ_CHECK(t0 < a.n_tiles && (kQ > 1 || a.n_keys >= 0));
    if (!uniform) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const unsigned g = unit_lo + warp + j * S::kWarps;
            if (g < unit_hi) v[j] = ld_stream_u32x8(lane_fp + static_cast<size_t>(g) * kFpPerUnit);
        }
    }
    if constexpr (kQ > 1) {
        if (tid < kQ) {
            int nk = 0;
            if (tid < nq) {
                if constexpr (kParamQuery) nk = min(sq.n_keys[tid], kP);
                else nk = min(a.n_keys_g[tid], S::kKeys);
            }
            sm.n_keys[tid] = nk;
        }
    }
    __syncthreads();
    mark(2);
    if constexpr (kQ == 1) {
        for (int i = tid; i < a.n_keys; i += S::kThreads) {
            unsigned long long k;
            int m;
            if constexpr (kParamQuery) { k = sq.keys[0][i]; m = sq.mult[0][i]; }
            else { k = a.keys_g[i]; m = resident ? a.mult_g[i] : 0; }
            if (resident) {
                sm.keys[0][i] = k;
                sm.mult[0][i] = m;
            }
            sm.map[filter_hash(k)] = 1;
        }
    } else {
        const int stride = kParamQuery ? kP : a.key_stride;
        for (int i = tid; i < nq * stride; i += S::kThreads) {
            const int b = i / stride, k = i - b * stride;
            if (k < sm.n_keys[b]) {
                unsigned long long key;
                int m;
                if constexpr (kParamQuery) { key = sq.keys[b][k]; m = sq.mult[b][k]; }
                else { key = a.keys_g[i]; m = a.mult_g[i]; }
                sm.keys[b][k] = key;
                sm.mult[b][k] = m;
                const uint32_t h = filter_hash(key);   // byte-wide OR through the containing 32-bit word
                atomicOr(reinterpret_cast<unsigned *>(sm.map) + (h >> 2), (1u << b) << (8 * (h & 3)));
            }
        }
    }
    // the compaction will read the video ids of this tile's rows: on their way to L2 now
    if (tid * 32 < td.n_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.vid + td.row_lo + tid * 32));
    __syncthreads();
    mark(3);
    pdl_wait();               // the previous query on this workspace still owns state / out until here
    pdl_launch_dependents();  // the next kernel may run its own prologue while this one streams
    mark(4);

    // ---- stream the tile's fingerprints; survivors are parked per warp and verified at the end ----
    const long long pos0 = static_cast<long long>(unit_lo) * kFpPerUnit;
    auto resolve = [&](unsigned rel) {
        TVZ_CHECK(pos0 + rel < a.n_pos && pos0 + rel < static_cast<long long>(unit_hi) * kFpPerUnit);
        const uint4 r4 = __ldg(reinterpret_cast<const uint4 *>(a.rec + pos0 + rel));
        const unsigned long long val = (static_cast<unsigned long long>(r4.y) << 32) | r4.x;
        const unsigned local = r4.z - static_cast<unsigned>(td.row_lo);
        if (local >= static_cast<unsigned>(td.n_rows)) return;   // a neighbour's row in a shared boundary unit
        if (kQ == 1) {
            int m;
            if (resident) {
                const int lo = lower_bound_u64([&](int i) { return sm.keys[0][i]; }, a.n_keys, val);
                if (lo >= a.n_keys || sm.keys[0][lo] != val) return;   // fingerprint collision (or padding)
                m = sm.mult[0][lo];
            } else {
                const int lo = lower_bound_u64([&](int i) { return __ldg(a.keys_g + i); }, a.n_keys, val);
                if (lo >= a.n_keys || __ldg(a.keys_g + lo) != val) return;
                m = __ldg(a.mult_g + lo);
            }
            TVZ_CHECK(local < static_cast<unsigned>(S::kRows));
            atomicAdd(&sm.counts[local], static_cast<unsigned>(m));
        } else {
            unsigned qm = sm.map[filter_hash(val)];
            while (qm) {
                const int b = __ffs(qm) - 1;
                qm &= qm - 1;
                const int nk = sm.n_keys[b];
                const int lo = lower_bound_u64([&](int i) { return sm.keys[b][i]; }, nk, val);
                if (lo >= nk || sm.keys[b][lo] != val) continue;       // filter false positive for this query
                TVZ_CHECK(b < kQ && local < static_cast<unsigned>(S::kRows));
                atomicAdd(&sm.counts[b * (S::kRows / 2) + (local >> 1)],
                          static_cast<unsigned>(sm.mult[b][lo]) << (16 * (local & 1)));
            }
        }
    };
    unsigned *qe = sm.qe[warp];
    int queued = 0;  // warp-uniform, <= kQueue
    auto drain = [&]() {
        __syncwarp();
        for (int i = lane; i < queued; i += 32) resolve(qe[i]);
        __syncwarp();
        queued = 0;
    };
    for (unsigned g0 = unit_lo + warp; g0 < unit_hi; g0 += 2 * S::kWarps) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const unsigned g = g0 + j * S::kWarps;
            if (g >= unit_hi) break;  // warp-uniform
            unsigned flags = 0;       // bit k: the lane's k-th fingerprint is in the byte map
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned w = v[j].w[k];
                if (kQ == 1) {   // the map holds 0 / 1
                    flags |= static_cast<unsigned>(sm.map[w & 0xffffu]) << (2 * k);
                    flags |= static_cast<unsigned>(sm.map[w >> 16]) << (2 * k + 1);
                } else {         // ... or one bit per query
                    flags |= static_cast<unsigned>(sm.map[w & 0xffffu] != 0) << (2 * k);
                    flags |= static_cast<unsigned>(sm.map[w >> 16] != 0) << (2 * k + 1);
                }
            }
            // park the survivors' positions, one per lane and round (a lane rarely holds two):
            // slots from the vote mask, no atomics, no scan
            unsigned mask = __ballot_sync(0xffffffffu, flags != 0u);
            if (mask) {
                const unsigned e0 = (g - unit_lo) * kFpPerUnit + lane * 16;
                do {
                    if (queued + __popc(mask) > S::kQueue) drain();
                    if (flags) {
                        const unsigned e = e0 + (__ffs(flags) - 1);
                        TVZ_CHECK(queued + __popc(mask & ((1u << lane) - 1u)) < S::kQueue);
                        qe[queued + __popc(mask & ((1u << lane) - 1u))] = e;
                        flags &= flags - 1;
                        // what the verification will read, on its way to L2 while the stream goes on
                        warm_record(a.rec + pos0 + e);
                    }
                    queued += __popc(mask);
                    mask = __ballot_sync(0xffffffffu, flags != 0u);
                } while (mask);
            }
            const unsigned gn = g + 2 * S::kWarps;
            if (gn < unit_hi) v[j] = ld_stream_u32x8(lane_fp + static_cast<size_t>(gn) * kFpPerUnit);
        }
    }
    mark(5);
    drain();
    __syncthreads();   // every count of this tile is final
    mark(6);

    // ---- compaction of the tile's own rows ----
    const int r0 = tid * S::kRowsPerThread;
    const bool want_all = a.min_match <= 0;   // then rows without any match qualify too -- except replaced ones
    auto count_of = [&](int b, int local) -> int {
        if (kQ == 1) return static_cast<int>(sm.counts[local]);
        return static_cast<int>((sm.counts[b * (S::kRows / 2) + (local >> 1)] >> (16 * (local & 1))) & 0xffffu);
    };
    auto qualifies = [&](int c, int local) -> bool {
        if (local >= td.n_rows || c < a.min_match) return false;
        return !(want_all && a.dead && a.dead[td.row_lo + local]);
    };
    for (int b = 0; b < nq; ++b) {
        int mine = 0;
#pragma unroll
        for (int j = 0; j < S::kRowsPerThread; ++j) mine += qualifies(count_of(b, r0 + j), r0 + j);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0) sm.warp_tot[b][warp] = mine;
    }
    __syncthreads();
    // Tile totals are exchanged through L2: every tile publishes {query sequence, total}; every thread then
    // polls its share of the PREDECESSORS' entries (relaxed loads: no fence, no L1 invalidation per poll)
    // until they carry this query's sequence -- one round trip once they are there, no look-back chain.
    // Only earlier tiles are waited for: they belong to CTAs with a lower blockIdx, which the hardware
    // dispatches first, so a CTA never waits for one that is not running.
    if (warp < nq) {   // warp b: exclusive prefix of the warps' totals of query b; publish the tile's total
        const int b = warp;
        const unsigned w = lane < S::kWarps ? sm.warp_tot[b][lane] : 0u;
        unsigned incl = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane < S::kWarps) sm.warp_tot[b][lane] = incl - w;
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);   // <= S::kRows <= 8192: fits 16 bits
        if (lane == 0) {
            sm.agg[b] = total;
            st_relaxed_u32(&a.state[static_cast<size_t>(tile) * kQ + b], (a.seq << 16) | total);
        }
    }
    mark(7);
    // A tile without a single qualifying row (most tiles of a selective query) has nothing to place: it
    // needs no offset, polls nothing and is done -- only the last tile always goes on (it writes the header).
    __syncthreads();
    bool idle = !last_cta;
    for (int b = 0; b < nq; ++b) idle = idle && sm.agg[b] == 0;
    if (!idle) {   // hits of all earlier tiles
        const unsigned seq = a.seq;
        const int b = tid / S::kGroup, i0 = tid - b * S::kGroup;
        unsigned long long sum = 0;
        if (b < nq) {
            for (int i = i0; i < tile; i += S::kGroup) {
                const unsigned *p = &a.state[static_cast<size_t>(i) * kQ + b];
                unsigned r = ld_poll_u32(p), polls = 0;
                while ((r >> 16) != seq) {   // bounded: a protocol bug must surface as a launch failure, never as a hung GPU
                    if (++polls == (1u << 24)) __trap();
                    r = ld_poll_u32(p);
                }
                sum += r & 0xffffu;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) sm.part[warp] = sum;
        __syncthreads();
        mark(8);
        if (tid < nq) {
            constexpr int kWarpsPerGroup = S::kGroup / 32;
            unsigned long long e = 0;
            for (int w = 0; w < kWarpsPerGroup; ++w) e += sm.part[tid * kWarpsPerGroup + w];
            sm.excl[tid] = e;
            if (last_cta) {   // the last tile knows every query's total
                const long long total = static_cast<long long>(e) + sm.agg[tid];
                a.n_hits_out[tid] = total;
                int *o = a.out + tid * a.out_stride;
                const int2 hdr = make_int2(total > 0x7fffffffll ? 0x7fffffff : static_cast<int>(total), total > a.cap ? 1 : 0);
                *reinterpret_cast<int2 *>(o) = hdr;
                for (int p = 0; p < a.gt.n_dst; ++p)   // fused gather: the header travels like any hit, as entry 0
                    st_tagged(a.gt.record[p] + tid * a.gt.query_stride, hdr, a.gt.epoch);
            }
        }
        __syncthreads();
        mark(9);
        if constexpr (TVZ_STAGED_EMIT == 2 || (TVZ_STAGED_EMIT == 1 && kQ > 1)) {
        // The hits go through shared memory -- the byte map's 64 KB are free once the stream has ended -- so that
        // consecutive threads write consecutive entries: a warp's stores to the record (and, fused gather, its tagged
        // 16-byte stores to the peers) cover contiguous 256 (512) bytes instead of 32 scattered entries.
        int2 *stage = reinterpret_cast<int2 *>(sm.map);
        static_assert(S::kRows * sizeof(int2) <= kMapEntries, "the staging area is the byte map");
        for (int b = 0; b < nq; ++b) {
            int cnt[S::kRowsPerThread];
            int mine = 0;
#pragma unroll
            for (int j = 0; j < S::kRowsPerThread; ++j) {
                cnt[j] = count_of(b, r0 + j);
                if (qualifies(cnt[j], r0 + j)) ++mine; else cnt[j] = -1;   // counts are never negative
            }
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            const long long base = static_cast<long long>(sm.excl[b]);
            int lpos = static_cast<int>(sm.warp_tot[b][warp]) + (incl - mine);   // rank of the thread's first hit within the CTA
#pragma unroll
            for (int j = 0; j < S::kRowsPerThread; ++j) {
                if (cnt[j] >= 0) {
                    const int row = td.row_lo + r0 + j;
                    TVZ_CHECK(row < a.n_rows_cap && r0 + j < td.n_rows && lpos >= 0 && lpos < S::kRows);
                    stage[lpos] = make_int2(a.vid[row], cnt[j]);
                    if (kQ == 1 && a.rows_out && base + lpos < a.cap) a.rows_out[base + lpos] = row;
                    ++lpos;
                }
            }
            __syncthreads();
            const int total = static_cast<int>(sm.agg[b]);
            int *o = a.out + b * a.out_stride;
            for (int i = tid; i < total; i += S::kThreads) {
                const long long pos = base + i;
                if (pos < a.cap) {
                    const int2 hit = stage[i];
                    *reinterpret_cast<int2 *>(o + 2 + 2 * pos) = hit;
                    // fused gather: every CTA ships its own hits to all peers as it places them (NVLink stores, in
                    // parallel across CTAs), each 8-byte half tagged with the query epoch
                    for (int p = 0; p < a.gt.n_dst; ++p)
                        st_tagged(a.gt.record[p] + b * a.gt.query_stride + 4 * (1 + pos), hit, a.gt.epoch);
                }
            }
            if (b + 1 < nq) __syncthreads();   // the next query's hits reuse the staging area
        }
        } else {
        for (int b = 0; b < nq; ++b) {
            int cnt[S::kRowsPerThread];
            int mine = 0;
#pragma unroll
            for (int j = 0; j < S::kRowsPerThread; ++j) {
                cnt[j] = count_of(b, r0 + j);
                if (qualifies(cnt[j], r0 + j)) ++mine; else cnt[j] = -1;   // counts are never negative
            }
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            long long pos = static_cast<long long>(sm.excl[b]) + sm.warp_tot[b][warp] + (incl - mine);
            int *o = a.out + b * a.out_stride;
#pragma unroll
            for (int j = 0; j < S::kRowsPerThread; ++j) {
                if (cnt[j] >= 0) {
                    TVZ_CHECK(pos >= 0);
                    if (pos < a.cap) {
                        const int row = td.row_lo + r0 + j;
                        TVZ_CHECK(row < a.n_rows_cap && r0 + j < td.n_rows);
                        const int2 hit = make_int2(a.vid[row], cnt[j]);
                        *reinterpret_cast<int2 *>(o + 2 + 2 * pos) = hit;
                        if (kQ == 1 && a.rows_out) a.rows_out[pos] = row;
                        // fused gather: every CTA ships its own hits to all peers as it finds them (NVLink stores, in
                        // parallel across CTAs), each 8-byte half tagged with the query epoch
                        for (int p = 0; p < a.gt.n_dst; ++p)
                            st_tagged(a.gt.record[p] + b * a.gt.query_stride + 4 * (1 + pos), hit, a.gt.epoch);
                    }
                    ++pos;
                }
            }
        }
        }
    }   // !idle
    mark(10);

    // ---- fused gather: the last CTAs wait until every peer's record for this query is complete here ----
    // Records travel in a tagged form (the LL idea of collective libraries): entry e of a record is 16 bytes
    // {value0, epoch, value1, epoch} -- entry 0 = {n_hits, overflow}, entry 1 + h = {video_id, match_count} -- and
    // each 8-byte half is stored atomically, so a reader that sees the epoch in both halves has the data.  Senders
    // therefore need no system-scope fence, no done-counter and no flag: CTAs just store and exit.  The receivers
    // are the LAST kRecvCtas CTAs of the grid.  Each polls the headers of all n_peers x n_queries lists in this
    // rank's OWN memory (one thread per header, side by side), then the lists are swept as ONE index space cut
    // into equal shares, a share per receiver (four independent 16-byte loads per thread and round; an entry that
    // is not there yet is polled, bounded) -- so the wait costs two round trips however the hits are spread over
    // peers and queries (one receiver per peer, the first form, took 17 rounds for 68 k hits per peer at N = 2).
    // When the kernel completes, all records are complete here.
    if (a.gt.n_peers == 0) return;
    const int back = static_cast<int>(gridDim.x) - 1 - tile;            // 0 for the last CTA
    const int n_recv = min(static_cast<int>(gridDim.x), kRecvCtas);
    if (back >= n_recv) return;
    mark(11);
    const int n_lists = a.gt.n_peers * nq;                              // <= kMaxPeers * kBatch
    long long *first = reinterpret_cast<long long *>(&sm.qe[0][0]);     // [n_lists + 1]; the survivor queues are long done
    static_assert(sizeof(sm.qe) >= (kMaxPeers * kBatch + 1) * sizeof(long long), "list offsets live in the survivor queues");
    auto list_slot = [&](int l) -> const int * {
        const int p = l / nq, b = l - p * nq;
        return a.gt.my_slots + p * a.gt.slot_stride + b * a.gt.query_stride;
    };
    if (tid < n_lists) first[1 + tid] = min(static_cast<long long>(ld_tagged(list_slot(tid), a.gt.epoch).x), a.cap);
    if (tid == 0) first[0] = 0;
    __syncthreads();
    if (warp == 0) {   // lengths -> exclusive offsets
        long long carry = 0;
        for (int c = 0; c < n_lists; c += 32) {
            long long incl = c + lane < n_lists ? first[1 + c + lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            if (c + lane < n_lists) first[1 + c + lane] = carry + incl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    const long long n_all = first[n_lists];
    const long long share = (n_all + n_recv - 1) / n_recv;
    const long long f_lo = back * share, f_hi = min(n_all, f_lo + share);
    auto entry_of = [&](long long f) -> const int * {   // the list whose range [first[l], first[l + 1]) holds f
        int l = 0, h = n_lists;
        while (h - l > 1) {
            const int mid = (l + h) >> 1;
            if (first[mid] <= f) l = mid; else h = mid;
        }
        return list_slot(l) + 4 * (1 + (f - first[l]));
    };
    for (long long f0 = f_lo + tid; f0 < f_hi; f0 += 4 * S::kThreads) {
        const int *e[4];
        bool there[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long f = f0 + u * S::kThreads;
            e[u] = f < f_hi ? entry_of(f) : nullptr;
            there[u] = !e[u] || try_tagged(e[u], a.gt.epoch);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (!there[u]) (void)ld_tagged(e[u], a.gt.epoch);
    }
    mark(15);
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

// Row upsert on the device (db.py:43-64): neutralise the replaced row, append the new one to the tail.
struct UpsertArgs {
    VerifyRec *rec;
    unsigned short *fp;
    unsigned long long *ts;
    long long *off;
    int *vid;
    unsigned char *dead;
    long long kill_row;                    // -1: nothing to replace
    long long kill_pos_lo, kill_pos_hi;    // arranged positions that may hold the replaced row's records
    long long kill_ts_lo, kill_ts_hi;      // its values in row order
    long long new_row, pos0, ts0;          // where the new row goes: arranged position (tail: identity order) and row-order position
    long long n_pos, n_ts, n_rows_cap;     // checked build: what the arrays hold
    int new_vid, n;
    const unsigned long long *vals_g;      // n > kParamKeys: values staged in device memory
    unsigned long long vals[kParamKeys];
};
__global__ void __launch_bounds__(512) upsert_kernel(const __grid_constant__ UpsertArgs u) {
    TVZ_CHECK(u.n >= 0 && u.pos0 >= 0 && u.pos0 + u.n <= u.n_pos && u.ts0 + u.n <= u.n_ts && u.new_row >= 0 && u.new_row < u.n_rows_cap);
    TVZ_CHECK(u.kill_row < u.n_rows_cap && (u.kill_row < 0 || (u.kill_pos_lo >= 0 && u.kill_pos_hi <= u.n_pos && u.kill_ts_hi <= u.n_ts)));
    if (u.kill_row >= 0) {
        for (long long p = u.kill_pos_lo + threadIdx.x; p < u.kill_pos_hi; p += blockDim.x)
            if (u.rec[p].row == static_cast<unsigned>(u.kill_row)) u.rec[p].ts = kPadPattern;
        for (long long p = u.kill_ts_lo + threadIdx.x; p < u.kill_ts_hi; p += blockDim.x) u.ts[p] = kPadPattern;
        if (threadIdx.x == 0 && u.kill_row != u.new_row) u.dead[u.kill_row] = 1;
    }
    __syncthreads();   // a row rewritten in place: its old records are gone before the new ones land
    for (int i = threadIdx.x; i < u.n; i += blockDim.x) {
        const unsigned long long v = u.vals_g ? u.vals_g[i] : u.vals[i];
        VerifyRec r;
        r.ts = v;
        r.row = static_cast<unsigned>(u.new_row);
        r.pad = 0;
        u.rec[u.pos0 + i] = r;
        u.fp[u.pos0 + i] = static_cast<unsigned short>(filter_hash(v));
        u.ts[u.ts0 + i] = v;
    }
    if (threadIdx.x == 0) {
        u.off[u.new_row] = u.ts0;
        u.off[u.new_row + 1] = u.ts0 + u.n;
        u.vid[u.new_row] = u.new_vid;
        u.dead[u.new_row] = 0;
    }
}

}  // namespace
}  // namespace tvz

using namespace tvz;

struct TailRow {
    int vid;
    long long start;   // first value, relative to the tail
    int n;
    bool alive;
};

struct tvz_catalog {
    int device = 0;
    long long n_rows_main = 0, n_vals_main = 0;
    long long n_units_main = 0;              // fingerprint units of the packed part
    long long ts_main_padded = 0;            // row-order values of the packed part (+ padding); the tail follows
    unsigned short *d_fp = nullptr;          // filter_hash of every stored value, padded to whole 512-value units and
                                             // arranged inside each unit for conflict-free lookups (arrange_fingerprints)
    VerifyRec *d_rec = nullptr;              // the value behind d_fp[p] and its row, in the same arranged order
    unsigned long long *d_ts = nullptr;
    long long *d_off = nullptr;
    int *d_vid = nullptr;
    unsigned char *d_dead = nullptr;
    TileDesc *d_tiles = nullptr;
    std::vector<TileDesc> tiles;             // tiles of the packed rows
    int uniform_units = 0;                   // > 0: packed tile t starts at unit t * uniform_units
    long long n_pos_all = 0, n_ts_all = 0, rows_cap = 0;   // what the device arrays hold (checked build)
    int tail_index = -1;
    // ---- mutable tail (tvz_catalog_upsert); everything below is guarded by `mu` ----
    mutable std::mutex mu;
    long long tail_cap_vals = 0;             // 0: immutable catalogue
    long long tail_used_vals = 0;
    int tail_used_rows = 0;
    std::vector<TailRow> tail_rows;
    std::vector<unsigned long long> tail_vals;        // host mirror of the tail's values (tail compaction)
    std::vector<long long> h_off;                     // packed rows: where a replaced row's values live
    std::unordered_map<int, long long> row_of_vid;    // first live row of a video (db.py:47 .first()); built on first upsert
    bool row_map_built = false;
    long long dead_main = 0;
    cudaStream_t mut_stream = nullptr;
    cudaEvent_t mut_event = nullptr;
    std::atomic<unsigned long long> mut_seq{0};
    unsigned long long *d_stage = nullptr;   // long rows: values staged for the upsert kernel
    unsigned long long *h_stage = nullptr;   // pinned
    long long stage_cap = 0;

    long long n_rows() const { return n_rows_main + tail_used_rows; }
    int tail_tiles() const { return tail_index < 0 ? 0 : std::max(1, (tail_used_rows + kTileRows - 1) / kTileRows); }
    long long n_tiles() const { return static_cast<long long>(tiles.size()) + tail_tiles(); }
    long long max_tiles() const { return static_cast<long long>(tiles.size()) + (tail_index < 0 ? 0 : kMaxTailTiles); }
};

struct tvz_match_ws {
    const tvz_catalog *cat = nullptr;
    long long cap = 0;
    unsigned *d_state = nullptr;            // [tiles][kBatch] {query sequence << 16 | qualifying rows of the tile}
    unsigned seq = 0;                       // sequence number of the last query enqueued (1..65535, 0 = none yet)
    bool wrapped_once = false;              // the sequence has been through 65535 at least once
    int *d_out = nullptr;                   // [cap+1][2]
    long long *d_rows = nullptr;            // [cap]
    int *d_kth = nullptr;                   // [cap]
    long long *d_nhits = nullptr;           // [kBatch]
    // query staging: q_canon u64 [q_cap] | keys u64 [q_cap] | mult i32 [q_cap]
    unsigned long long *d_keys = nullptr, *d_qcanon = nullptr;
    int *d_mult = nullptr;
    uint8_t *h_stage = nullptr;    // pinned
    size_t stage_bytes = 0;
    int q_cap = 0;
    int *h_out = nullptr;          // pinned [cap+1][2]
    int *h_kth = nullptr;          // pinned [cap]
    cudaStream_t stream = nullptr; // private stream for the synchronous entry points
    cudaEvent_t staged = nullptr;  // the pinned staging buffer has been consumed
    bool stage_busy = false;
    bool timing = false;           // debug: bracket the kernel with events
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    long long *d_trace = nullptr;           // debug (not owned)
    int seen_tiles = 0;                     // largest grid a query of this workspace has launched
    unsigned long long seen_mut = 0;        // last catalogue mutation a stream of this workspace has waited for
    cudaStream_t seen_stream = nullptr;     // ... and which stream that was
    // batched queries (allocated on first use)
    int *b_out = nullptr;                   // [kBatch][cap+1][2]
    uint8_t *b_dev = nullptr;               // keys u64 [kBatch][kParamKeys] | mult i32 [..] | n_keys i32 [kBatch]
    uint8_t *hb_stage = nullptr;            // pinned: kStageSlots mirrors of b_dev, used round-robin
    cudaEvent_t b_staged[4] = {};           // slot i has been copied to the device
    bool b_stage_busy[4] = {};
    int b_slot = 0;
    int *hb_out = nullptr;                  // pinned [kBatch][cap+1][2]
};

namespace {

#if !TVZ_BATCH_PARAMS   // the staged form of a batch's keys (measurement variant, scripts/batch_variants.sh)
constexpr size_t kBatchKeysBytes = static_cast<size_t>(kBatch) * kParamKeys * 8;
constexpr size_t kBatchMultBytes = static_cast<size_t>(kBatch) * kParamKeys * 4;
constexpr size_t kBatchStageBytes = (kBatchKeysBytes + kBatchMultBytes + kBatch * 4 + 63) / 64 * 64;
constexpr int kStageSlots = 4;   // pinned staging slots: the host prepares batch k+1..k+3 while batch k still waits for its copy
#endif

inline unsigned long long canon_bits(double v) {
    unsigned long long b;
    memcpy(&b, &v, 8);
    if (b == 0x8000000000000000ull) b = 0;  // -0.0 == 0.0 in Python
    return b;
}
inline bool is_nan_bits(unsigned long long b) {
    return (b & 0x7ff0000000000000ull) == 0x7ff0000000000000ull && (b & 0x000fffffffffffffull) != 0;
}

// One row's stored form: NaN dropped, -0.0 folded, in-row repeats dropped (first occurrence kept).
void canon_row(const double *v, long long n, std::vector<unsigned long long> &out, std::unordered_set<unsigned long long> &seen) {
    const size_t start = out.size();
    bool ascending = true;
    double last = 0;
    for (long long j = 0; j < n; ++j) {
        const unsigned long long b = canon_bits(v[j]);
        if (is_nan_bits(b)) continue;
        if (out.size() > start && !(v[j] > last)) ascending = false;
        last = v[j];
        out.push_back(b);
    }
    if (!ascending) {  // rare: unsorted or repeated values -> stable de-duplication
        seen.clear();
        size_t w = start;
        for (size_t j = start; j < out.size(); ++j)
            if (seen.insert(out[j]).second) out[w++] = out[j];
        out.resize(w);
    }
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

// Cut rows [0, n_rows) into tiles of whole rows: at most kTileRows rows each, about the same number of
// stored values each, about `want` tiles (one wave of CTAs; a multiple of it when rows are short).
// Preferred form ("uniform", returns U > 0): tile t owns the rows whose first value lies in units
// [t * U, (t + 1) * U) -- its first unit is arithmetic, and it streams on to the end of its last row.
// Catalogues that cannot be cut that way with <= kTileRows rows per tile (e.g. thousands of empty rows in one
// place) fall back to a greedy cut with explicit unit ranges (returns 0).
int build_tiles(const std::vector<long long> &off, long long n_rows, int want, std::vector<TileDesc> &tiles) {
    tiles.clear();
    if (n_rows <= 0) return 0;
    const long long n_vals = off[n_rows];
    const long long n_units = (n_vals + kFpPerUnit - 1) / kFpPerUnit;
    const long long by_rows = (n_rows + kTileRows - 1) / kTileRows;
    const long long waves = std::max<long long>(1, (by_rows + want - 1) / want);
    long long U = std::max<long long>(kMinTileUnits, (n_units + waves * want - 1) / (waves * want));
    for (int attempt = 0; attempt < 24 && n_units > 0; ++attempt) {
        const long long T = (n_units + U - 1) / U;
        if (T > 16ll * want) break;
        const long long W = U * kFpPerUnit;
        tiles.clear();
        bool ok = true;
        long long r = 0;
        for (long long t = 0; t < T && ok; ++t) {
            // rows whose first value lies below (t + 1) * W; the last tile takes whatever is left (trailing empty rows)
            const long long e = t == T - 1 ? n_rows
                                           : std::lower_bound(off.begin() + r, off.begin() + n_rows, (t + 1) * W) - off.begin();
            if (e - r > kTileRows) { ok = false; break; }
            TileDesc d;
            d.row_lo = static_cast<int>(r);
            d.n_rows = static_cast<int>(e - r);
            d.unit_lo = static_cast<unsigned>(t * U);
            d.unit_hi = e > r && off[e] > off[r] ? static_cast<unsigned>((off[e] + kFpPerUnit - 1) / kFpPerUnit) : d.unit_lo;
            tiles.push_back(d);
            r = e;
        }
        if (ok) return static_cast<int>(U);
        if (U == kMinTileUnits) break;
        U = std::max<long long>(kMinTileUnits, U * 3 / 4);
    }
    // greedy fallback
    tiles.clear();
    const long long target = std::max<long long>(static_cast<long long>(kMinTileUnits) * kFpPerUnit,
                                                 (n_vals + waves * want - 1) / (waves * want));
    long long r = 0;
    while (r < n_rows) {
        // first row index e with off[e] - off[r] >= target
        long long e = std::lower_bound(off.begin() + r, off.begin() + n_rows + 1, off[r] + target) - off.begin();
        e = std::min<long long>({e, r + kTileRows, n_rows});
        if (e <= r) e = r + 1;
        if (n_rows - e < kTileRows / 8 && n_rows - r <= kTileRows && off[n_rows] - off[r] < target + target / 4) e = n_rows;  // no sliver at the end
        TileDesc t;
        t.row_lo = static_cast<int>(r);
        t.n_rows = static_cast<int>(e - r);
        t.unit_lo = static_cast<unsigned>(off[r] / kFpPerUnit);
        t.unit_hi = off[e] == off[r] ? t.unit_lo : static_cast<unsigned>((off[e] + kFpPerUnit - 1) / kFpPerUnit);
        tiles.push_back(t);
        r = e;
    }
    return 0;
}

// Tail tile k holds tail rows [k * 4096, (k + 1) * 4096) and the units their values span (tail rows are
// contiguous in append order).  Called with the lock held.
void tail_descs(const tvz_catalog *c, TileDesc *out) {
    const int nt = c->tail_tiles();
    for (int k = 0; k < nt; ++k) {
        const int r0 = k * kTileRows, r1 = std::min(c->tail_used_rows, r0 + kTileRows);
        TileDesc t;
        t.row_lo = static_cast<int>(c->n_rows_main) + r0;
        t.n_rows = std::max(0, r1 - r0);
        const long long v0 = r1 > r0 ? c->tail_rows[r0].start : 0;
        const long long v1 = r1 > r0 ? c->tail_rows[r1 - 1].start + c->tail_rows[r1 - 1].n : 0;
        t.unit_lo = static_cast<unsigned>(c->n_units_main + v0 / kFpPerUnit);
        t.unit_hi = v1 > v0 ? static_cast<unsigned>(c->n_units_main + (v1 + kFpPerUnit - 1) / kFpPerUnit) : t.unit_lo;
        out[k] = t;
    }
}

int ensure_query_capacity(tvz_match_ws *ws, int qn) {
    if (qn <= ws->q_cap) return TVZ_OK;
    int cap = std::max(256, ws->q_cap);
    while (cap < qn) cap *= 2;
    if (ws->stage_busy) { TVZ_CUDA(cudaEventSynchronize(ws->staged)); ws->stage_busy = false; }
    if (ws->d_qcanon) cudaFree(ws->d_qcanon);
    if (ws->d_mult) cudaFree(ws->d_mult);
    if (ws->h_stage) cudaFreeHost(ws->h_stage);
    ws->d_qcanon = nullptr;
    ws->d_keys = nullptr;
    ws->d_mult = nullptr;
    ws->h_stage = nullptr;
    ws->q_cap = 0;
    TVZ_CUDA(cudaMalloc(&ws->d_qcanon, sizeof(unsigned long long) * cap * 2));  // canonical query, then sorted keys
    ws->d_keys = ws->d_qcanon + cap;
    TVZ_CUDA(cudaMalloc(&ws->d_mult, sizeof(int) * cap));
    ws->stage_bytes = sizeof(unsigned long long) * cap * 2 + sizeof(int) * cap + 64;
    TVZ_CUDA(cudaHostAlloc(&ws->h_stage, ws->stage_bytes, cudaHostAllocDefault));
    ws->q_cap = cap;
    return TVZ_OK;
}

static_assert(sizeof(TileSmem<1>) <= 227 * 1024 && sizeof(TileSmem<kBatch>) <= 227 * 1024, "a CTA has 227 KB of shared memory");
static_assert(sizeof(TileArgs) + sizeof(QueryParam<kBatch, kParamKeys>) <= 32764, "kernel parameters are limited to 32,764 bytes");
template <int kQ, int kP>
cudaError_t set_tile_attr() {
    return cudaFuncSetAttribute(match_tile_kernel<kQ, kP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(sizeof(TileSmem<kQ>)));
}
// Function attributes are set once per device, not per launch.
int ensure_kernel_attrs() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    TVZ_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return TVZ_OK;
    TVZ_CUDA((set_tile_attr<1, kParamKeys>()));
    TVZ_CUDA((set_tile_attr<1, 0>()));
#if TVZ_BATCH_PARAMS
    TVZ_CUDA((set_tile_attr<kBatch, kBatchShortKeys>()));
    TVZ_CUDA((set_tile_attr<kBatch, kParamKeys>()));
#else
    TVZ_CUDA((set_tile_attr<kBatch, 0>()));
#endif
    if (dev >= 0 && dev < 64) done[dev] = true;
    return TVZ_OK;
}

// The catalogue as a query sees it: tail tile and mutation sequence, read under the lock.
struct CatView {
    TileDesc tail[kMaxTailTiles] = {};
    int n_tiles = 0;
    unsigned long long seq = 0;
};
CatView view_of(const tvz_catalog *cat) {
    CatView v;
    if (cat->tail_index >= 0) {
        std::lock_guard<std::mutex> lk(cat->mu);
        tail_descs(cat, v.tail);
        v.n_tiles = static_cast<int>(cat->n_tiles());
        v.seq = cat->mut_seq.load(std::memory_order_relaxed);
    } else {
        v.n_tiles = static_cast<int>(cat->tiles.size());
    }
    return v;
}

void base_args(const tvz_catalog *cat, tvz_match_ws *ws, const CatView &cv, TileArgs &a) {
    a.fp = cat->d_fp;
    a.rec = cat->d_rec;
    a.tiles = cat->d_tiles;
    memcpy(a.tail, cv.tail, sizeof(a.tail));
    a.n_tiles = cv.n_tiles;
    a.tail_index = cat->tail_index;
    a.vid = cat->d_vid;
    a.dead = cat->d_dead;
    a.state = ws->d_state;
    a.n_hits_out = ws->d_nhits;
    a.uniform_units = cat->uniform_units;
    a.n_pos = cat->n_pos_all;
    a.n_rows_cap = cat->rows_cap;
    a.trace = ws->d_trace;
    ws->seq = ws->seq >= 0xffffu ? 1u : ws->seq + 1u;   // never 0: that is what fresh state[] entries carry
    a.seq = ws->seq;
}

// The 16-bit sequence has wrapped: an exchange entry that no query has rewritten for a whole cycle (slot b of a
// tile, after 65535 queries none of which was a batch with more than b queries) would look current again, so
// every entry goes back to "fresh" first.  Stream-ordered in front of the query that carries sequence 1.
int reset_state_on_wrap(const tvz_catalog *cat, tvz_match_ws *ws, cudaStream_t st) {
    if (ws->seq != 1u || !ws->wrapped_once) {
        ws->wrapped_once = ws->wrapped_once || ws->seq == 0xffffu;
        return TVZ_OK;
    }
    const size_t n_state = static_cast<size_t>(std::max<long long>(1, cat->max_tiles())) * kBatch;
    TVZ_CUDA(cudaMemsetAsync(ws->d_state, 0, n_state * 4, st));
    return TVZ_OK;
}

// A query must see every upsert that returned before it was enqueued.  (Also: tiles that appear for the
// first time -- the tail has grown -- start from untagged exchange entries.)
int wait_for_mutations(const tvz_catalog *cat, tvz_match_ws *ws, const CatView &cv, cudaStream_t st) {
    if (cv.n_tiles > ws->seen_tiles) {
        if (ws->seen_tiles > 0)
            TVZ_CUDA(cudaMemsetAsync(ws->d_state + static_cast<size_t>(ws->seen_tiles) * kBatch, 0,
                                     static_cast<size_t>(cv.n_tiles - ws->seen_tiles) * kBatch * 4, st));
        ws->seen_tiles = cv.n_tiles;
    }
    if (cat->tail_index >= 0 && cv.seq != 0 && (cv.seq != ws->seen_mut || st != ws->seen_stream)) {
        TVZ_CUDA(cudaStreamWaitEvent(st, cat->mut_event, 0));
        ws->seen_mut = cv.seq;
        ws->seen_stream = st;
    }
    return TVZ_OK;
}

// Enqueue query upload + the tile kernel (+ kth) on `st`.  `want_kth` needs qn <= q_cap.
int enqueue_match(const tvz_catalog *cat, tvz_match_ws *ws, const double *h_q, int qn, int min_match, bool want_kth,
                  int *d_out, long long out_cap, cudaStream_t st, const GatherTargets *gather = nullptr) {
    TVZ_REQUIRE(cat && ws && ws->cat == cat, "workspace does not belong to this catalogue");
    TVZ_REQUIRE(qn >= 0 && (qn == 0 || h_q), "bad query");
    DeviceGuard on_device(cat->device);
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
    int n_sorted = 0;
    bool ascending = true;   // by bit pattern, which is all the kernel's binary search needs
    for (int i = 0; i < qn; ++i) {
        const unsigned long long b = canon_bits(h_q[i]);
        h_qc[i] = b;
        if (is_nan_bits(b)) continue;
        if (n_sorted && b < h_keys[n_sorted - 1]) ascending = false;
        h_keys[n_sorted++] = b;
    }
    if (!ascending) std::sort(h_keys, h_keys + n_sorted);   // production queries are ascending cut lists (app.py:231)
    int nk = 0;
    for (int i = 0; i < n_sorted;) {
        int j = i;
        while (j < n_sorted && h_keys[j] == h_keys[i]) ++j;
        h_keys[nk] = h_keys[i];
        h_mult[nk] = j - i;
        ++nk;
        i = j;
    }
    // The pinned staging buffer is read by the device only when a copy out of it is enqueued; the
    // common case (a short query riding in the kernel parameters) never is, so back-to-back
    // asynchronous queries pipeline on the stream without a host-side wait in between.
    bool staged_copy = false;
    const CatView cv = view_of(cat);
    rc = wait_for_mutations(cat, ws, cv, st);
    if (rc) return rc;
    if (want_kth && qn > 0) {
        TVZ_CUDA(cudaMemcpyAsync(ws->d_qcanon, h_qc, sizeof(unsigned long long) * qn, cudaMemcpyHostToDevice, st));
        staged_copy = true;
    }
    if (cv.n_tiles > 0) {
        rc = ensure_kernel_attrs();
        if (rc) return rc;
        TileArgs a{};
        base_args(cat, ws, cv, a);
        rc = reset_state_on_wrap(cat, ws, st);
        if (rc) return rc;
        a.n_queries = 1;
        a.n_keys = nk;
        a.min_match = min_match;
        a.cap = out_cap;
        a.out = d_out;
        a.out_stride = 0;
        a.rows_out = ws->d_rows;
        if (gather) a.gt = *gather;
        const bool param = nk <= kParamKeys;
        if (!param) {
            TVZ_CUDA(cudaMemcpyAsync(ws->d_keys, h_keys, sizeof(unsigned long long) * nk, cudaMemcpyHostToDevice, st));
            TVZ_CUDA(cudaMemcpyAsync(ws->d_mult, h_mult, sizeof(int) * nk, cudaMemcpyHostToDevice, st));
            a.keys_g = ws->d_keys;
            a.mult_g = ws->d_mult;
            staged_copy = true;
        }
        if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t0, st));
        const dim3 grid(static_cast<unsigned>((cv.n_tiles + TileShape<1>::kPair - 1) / TileShape<1>::kPair)), block(TileShape<1>::kThreads);
        if (param) {
            QueryParam<1, kParamKeys> sq;
            memcpy(sq.keys[0], h_keys, sizeof(unsigned long long) * nk);
            memcpy(sq.mult[0], h_mult, sizeof(int) * nk);
            sq.n_keys[0] = nk;
            TVZ_CUDA(launch_pdl(match_tile_kernel<1, kParamKeys>, grid, block, sizeof(TileSmem<1>), st, a, sq));
        } else {
            TVZ_CUDA(launch_pdl(match_tile_kernel<1, 0>, grid, block, sizeof(TileSmem<1>), st, a, QueryParam<1, 0>{}));
        }
        if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t1, st));
        if (want_kth) {
            match_kth_kernel<<<2 * num_sms(), 256, 0, st>>>(cat->d_ts, cat->d_off, ws->d_rows, d_out, out_cap,
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

int ensure_batch_buffers(tvz_match_ws *ws) {
    if (ws->b_out) return TVZ_OK;
    TVZ_CUDA(cudaMalloc(&ws->b_out, kBatch * (ws->cap + 1) * 8));
    TVZ_CUDA(cudaHostAlloc(&ws->hb_out, kBatch * (ws->cap + 1) * 8, cudaHostAllocDefault));
#if !TVZ_BATCH_PARAMS
    TVZ_CUDA(cudaMalloc(&ws->b_dev, kBatchStageBytes));
    TVZ_CUDA(cudaHostAlloc(&ws->hb_stage, kStageSlots * kBatchStageBytes, cudaHostAllocDefault));
    for (int i = 0; i < kStageSlots; ++i) TVZ_CUDA(cudaEventCreateWithFlags(&ws->b_staged[i], cudaEventDisableTiming));
#endif
    return TVZ_OK;
}

// Sorted distinct values (canonical bit patterns) and multiplicities of query g0 + b into keys / mult [stride];
// -> number of distinct values.
int batch_query_keys(const double *q_all, const int64_t *q_off, int g, std::vector<unsigned long long> &sorted,
                     unsigned long long *keys, int *mult, int limit, int *nk_out) {
    const double *q = q_all + q_off[g];
    const long long qn = q_off[g + 1] - q_off[g];
    TVZ_REQUIRE(qn >= 0, "query offsets must be non-decreasing");
    TVZ_REQUIRE(qn <= 65535, "query %d has %lld values: not batchable (16-bit counts)", g, qn);
    sorted.clear();
    bool ascending = true;
    for (long long i = 0; i < qn; ++i) {
        const unsigned long long bits = canon_bits(q[i]);
        if (is_nan_bits(bits)) continue;
        if (!sorted.empty() && bits < sorted.back()) ascending = false;
        sorted.push_back(bits);
    }
    if (!ascending) std::sort(sorted.begin(), sorted.end());   // production queries are ascending cut lists (app.py:231)
    int nk = 0;
    for (size_t i = 0; i < sorted.size();) {
        size_t j = i;
        while (j < sorted.size() && sorted[j] == sorted[i]) ++j;
        if (nk < limit) {
            keys[nk] = sorted[i];
            mult[nk] = static_cast<int>(j - i);
        }
        ++nk;
        i = j;
    }
    *nk_out = nk;
    return TVZ_OK;
}

template <int kP>
int launch_batch_params(const tvz_match_ws *ws, const TileArgs &a, const unsigned long long *keys, const int *mult,
                        const int *nk, int nb, dim3 grid, cudaStream_t st) {
    QueryParam<kBatch, kP> qp;
    for (int b = 0; b < kBatch; ++b) {
        qp.n_keys[b] = b < nb ? nk[b] : 0;
        if (b < nb && nk[b] > 0) {
            memcpy(qp.keys[b], keys + static_cast<size_t>(b) * kParamKeys, sizeof(unsigned long long) * nk[b]);
            memcpy(qp.mult[b], mult + static_cast<size_t>(b) * kParamKeys, sizeof(int) * nk[b]);
        }
    }
    if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t0, st));
    TVZ_CUDA(launch_pdl(match_tile_kernel<kBatch, kP>, grid, dim3(TileShape<kBatch>::kThreads), sizeof(TileSmem<kBatch>), st, a, qp));
    if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t1, st));
    return TVZ_OK;
}

// Up to kBatch queries in one pass, one kernel; records land in d_out [nb][out_stride] (NULL: the workspace's
// own b_out with out_stride = 2 * (cap + 1)).  The keys ride in the kernel parameters (TVZ_BATCH_PARAMS; else:
// staged through pinned memory with ONE copy in front of the kernel).
int enqueue_batch(const tvz_catalog *cat, tvz_match_ws *ws, const double *q_all, const int64_t *q_off, int g0, int nb,
                  int min_match, int *d_out, long long out_cap, cudaStream_t st, const GatherTargets *gather = nullptr) {
    TVZ_REQUIRE(nb >= 1 && nb <= kBatch, "a batch holds 1..%d queries", kBatch);
    DeviceGuard on_device(cat->device);
    int rc = ensure_batch_buffers(ws);
    if (rc) return rc;
    if (!d_out) {
        d_out = ws->b_out;
        if (out_cap <= 0) out_cap = ws->cap;
    }
    TVZ_REQUIRE(out_cap >= 1 && out_cap <= ws->cap, "output capacity %lld outside [1, %lld]", out_cap, ws->cap);
#if TVZ_BATCH_PARAMS
    // thread-local scratch: concurrent callers have their own workspaces, but why allocate per call
    static thread_local std::vector<unsigned long long> keys_buf(static_cast<size_t>(kBatch) * kParamKeys);
    static thread_local std::vector<int> mult_buf(static_cast<size_t>(kBatch) * kParamKeys);
    unsigned long long *h_keys = keys_buf.data();
    int *h_mult = mult_buf.data();
    int nk_arr[kBatch] = {};
    int *h_nk = nk_arr;
#else
    const int slot = ws->b_slot;
    ws->b_slot = (slot + 1) % kStageSlots;
    if (ws->b_stage_busy[slot]) { TVZ_CUDA(cudaEventSynchronize(ws->b_staged[slot])); ws->b_stage_busy[slot] = false; }
    uint8_t *stage = ws->hb_stage + slot * kBatchStageBytes;
    unsigned long long *h_keys = reinterpret_cast<unsigned long long *>(stage);
    int *h_mult = reinterpret_cast<int *>(stage + kBatchKeysBytes);
    int *h_nk = reinterpret_cast<int *>(stage + kBatchKeysBytes + kBatchMultBytes);
    for (int b = 0; b < kBatch; ++b) h_nk[b] = 0;
#endif
    static thread_local std::vector<unsigned long long> sorted;
    int nk_max = 0;
    for (int b = 0; b < nb; ++b) {
        rc = batch_query_keys(q_all, q_off, g0 + b, sorted, h_keys + static_cast<size_t>(b) * kParamKeys,
                              h_mult + static_cast<size_t>(b) * kParamKeys, kParamKeys, &h_nk[b]);
        if (rc) return rc;
        TVZ_REQUIRE(h_nk[b] <= kParamKeys, "query %d has more than %d distinct values: not batchable", g0 + b, kParamKeys);
        nk_max = std::max(nk_max, h_nk[b]);
    }
    const CatView cv = view_of(cat);
    rc = wait_for_mutations(cat, ws, cv, st);
    if (rc) return rc;
    if (cv.n_tiles == 0) {
        TVZ_REQUIRE(!gather || gather->n_peers == 0, "an empty shard cannot take part in the fused gather");
        for (int b = 0; b < nb; ++b) TVZ_CUDA(cudaMemsetAsync(d_out + b * 2 * (out_cap + 1), 0, 8, st));
        TVZ_CUDA(cudaMemsetAsync(ws->d_nhits, 0, 8 * kBatch, st));
        return TVZ_OK;
    }
    rc = ensure_kernel_attrs();
    if (rc) return rc;
    TileArgs a{};
    base_args(cat, ws, cv, a);
    rc = reset_state_on_wrap(cat, ws, st);
    if (rc) return rc;
    a.key_stride = kParamKeys;
    a.n_queries = nb;
    a.min_match = min_match;
    a.cap = out_cap;
    a.out = d_out;
    a.out_stride = 2 * (out_cap + 1);
    a.rows_out = nullptr;
    if (gather) {
        a.gt = *gather;
        a.gt.query_stride = 4 * (out_cap + 1);   // tagged entries are 4 ints
    }
    const dim3 grid(static_cast<unsigned>((cv.n_tiles + TileShape<kBatch>::kPair - 1) / TileShape<kBatch>::kPair));
#if TVZ_BATCH_PARAMS
    if (nk_max <= kBatchShortKeys) return launch_batch_params<kBatchShortKeys>(ws, a, h_keys, h_mult, h_nk, nb, grid, st);
    return launch_batch_params<kParamKeys>(ws, a, h_keys, h_mult, h_nk, nb, grid, st);
#else
    TVZ_CUDA(cudaMemcpyAsync(ws->b_dev, stage, kBatchStageBytes, cudaMemcpyHostToDevice, st));
    TVZ_CUDA(cudaEventRecord(ws->b_staged[slot], st));
    ws->b_stage_busy[slot] = true;
    a.keys_g = reinterpret_cast<const unsigned long long *>(ws->b_dev);
    a.mult_g = reinterpret_cast<const int *>(ws->b_dev + kBatchKeysBytes);
    a.n_keys_g = reinterpret_cast<const int *>(ws->b_dev + kBatchKeysBytes + kBatchMultBytes);
    if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t0, st));
    TVZ_CUDA(launch_pdl(match_tile_kernel<kBatch, 0>, grid, dim3(TileShape<kBatch>::kThreads), sizeof(TileSmem<kBatch>), st,
                        a, QueryParam<kBatch, 0>{}));
    if (ws->timing) TVZ_CUDA(cudaEventRecord(ws->t1, st));
    return TVZ_OK;
#endif
}

int make_gather(int n_peers, int n_dst, const uint64_t *peer_record, const int32_t *d_my_slots, int64_t slot_stride_ints,
                uint32_t epoch, GatherTargets &gt) {
    TVZ_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers, "n_peers %d outside [1, %d]", n_peers, kMaxPeers);
    TVZ_REQUIRE(n_dst == n_peers || n_dst == 1, "peer_record holds one address per peer, or one multicast address");
    TVZ_REQUIRE(peer_record && d_my_slots && slot_stride_ints > 0 && slot_stride_ints % 4 == 0, "bad gather buffers");
    TVZ_REQUIRE(epoch != 0, "epoch 0 is what fresh gather buffers carry");
    gt.n_peers = n_peers;
    gt.n_dst = n_dst;
    gt.epoch = epoch;
    gt.my_slots = d_my_slots;
    gt.slot_stride = slot_stride_ints;
    for (int p = 0; p < n_dst; ++p) {
        TVZ_REQUIRE(peer_record[p] % 16 == 0, "gather slots must be 16-byte aligned");
        gt.record[p] = reinterpret_cast<int *>(static_cast<uintptr_t>(peer_record[p]));
    }
    return TVZ_OK;
}

void free_catalog_device(tvz_catalog *c) {
    void *dev[] = {c->d_fp, c->d_rec, c->d_ts, c->d_off, c->d_vid, c->d_dead, c->d_tiles, c->d_stage};
    for (void *p : dev)
        if (p) cudaFree(p);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->mut_event) cudaEventDestroy(c->mut_event);
    if (c->mut_stream) cudaStreamDestroy(c->mut_stream);
}

int catalog_create(const double *h_ts, const int64_t *h_off, const int32_t *h_video_id, int64_t n_rows,
                   int64_t tail_values, tvz_catalog **out) {
    TVZ_REQUIRE(out, "null out pointer");
    *out = nullptr;
    TVZ_REQUIRE(n_rows >= 0 && tail_values >= 0, "negative size");
    TVZ_REQUIRE(n_rows < (1ll << 31) - kTailRows, "too many rows for one shard");
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
        canon_row(h_ts + h_off[r], h_off[r + 1] - h_off[r], ts, seen);
        off[r + 1] = static_cast<long long>(ts.size());
    }
    TVZ_REQUIRE(static_cast<long long>(ts.size() / kFpPerUnit) + tail_values / kFpPerUnit + 4 < (1ll << 22),
                "too many stored values for one shard");
    tvz_catalog *c = new tvz_catalog();
    c->n_rows_main = n_rows;
    c->n_vals_main = static_cast<long long>(ts.size());
    c->n_units_main = (c->n_vals_main + kFpPerUnit - 1) / kFpPerUnit;
    c->ts_main_padded = c->n_units_main * kFpPerUnit;
    const bool is_mutable = tail_values > 0;
    c->tail_cap_vals = is_mutable ? (tail_values + kFpPerUnit - 1) / kFpPerUnit * kFpPerUnit : 0;
    const long long n_units_all = c->n_units_main + c->tail_cap_vals / kFpPerUnit;
    const long long rows_cap = n_rows + (is_mutable ? kTailRows : 0);

    // one wave of CTAs (2 per SM), one of them kept for the tail tile
    const int want = std::max(1, 2 * num_sms() - (is_mutable ? 1 : 0));
    c->uniform_units = build_tiles(off, n_rows, want, c->tiles);
    if (is_mutable) {
        c->tail_index = static_cast<int>(c->tiles.size());
        c->h_off = off;
    }

    // 16-bit fingerprints, padded to whole warp units (pad entries point past n_vals and are dropped)
    const size_t n_pos_main = static_cast<size_t>(std::max<long long>(1, c->n_units_main)) * kFpPerUnit;
    std::vector<unsigned short> fp(n_pos_main, 0);
    std::vector<unsigned short> perm(n_pos_main, 0);
    arrange_fingerprints(ts.data(), c->n_vals_main, c->n_units_main, fp.data(), perm.data());
    // verification records in ARRANGED order: a surviving fingerprint at position p is checked against
    // rec[p].ts and, if it is a real match, adds into the count of rec[p].row -- one 16-byte load, no search
    // for the row.  Pad positions hold a NaN pattern that equals no query key.
    std::vector<VerifyRec> rec(n_pos_main, VerifyRec{kPadPattern, 0u, 0u});
    {
        std::vector<unsigned> row_of(kFpPerUnit);
        long long r = 0;
        for (long long u = 0; u < c->n_units_main; ++u) {
            const long long base = u * kFpPerUnit;
            for (int i = 0; i < kFpPerUnit && base + i < c->n_vals_main; ++i) {
                while (r + 1 < n_rows && off[r + 1] <= base + i) ++r;
                row_of[i] = static_cast<unsigned>(r);
            }
            for (int p2 = 0; p2 < kFpPerUnit; ++p2) {
                const long long elem = base + perm[base + p2];
                if (elem < c->n_vals_main) {
                    rec[base + p2].ts = ts[elem];
                    rec[base + p2].row = row_of[perm[base + p2]];
                }
            }
        }
    }
    ts.resize(static_cast<size_t>(c->ts_main_padded), kPadPattern);

    cudaGetDevice(&c->device);
    auto fail = [&](cudaError_t e, const char *what) {
        set_error(TVZ_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        tvz_catalog_destroy(c);
        return TVZ_ERR_CUDA;
    };
    cudaError_t e;
    const size_t n_pos_all = static_cast<size_t>(std::max<long long>(1, n_units_all)) * kFpPerUnit;
    const size_t n_fp_alloc = n_pos_all + static_cast<size_t>(kFpPadUnits) * kFpPerUnit;
    const size_t n_ts_all = static_cast<size_t>(std::max<long long>(1, c->ts_main_padded + c->tail_cap_vals));
    c->n_pos_all = static_cast<long long>(n_pos_all);
    c->n_ts_all = static_cast<long long>(n_ts_all);
    c->rows_cap = rows_cap;
    if ((e = cudaMalloc(&c->d_fp, n_fp_alloc * 2)) != cudaSuccess) return fail(e, "cudaMalloc(fp)");
    if ((e = cudaMemset(c->d_fp, 0, n_fp_alloc * 2)) != cudaSuccess) return fail(e, "cudaMemset(fp)");
    if ((e = cudaMalloc(&c->d_rec, n_pos_all * sizeof(VerifyRec))) != cudaSuccess) return fail(e, "cudaMalloc(rec)");
    if ((e = cudaMalloc(&c->d_ts, n_ts_all * 8)) != cudaSuccess) return fail(e, "cudaMalloc(ts)");
    if ((e = cudaMalloc(&c->d_off, (rows_cap + 1) * 8)) != cudaSuccess) return fail(e, "cudaMalloc(off)");
    if ((e = cudaMalloc(&c->d_vid, std::max<size_t>(1, rows_cap) * 4)) != cudaSuccess) return fail(e, "cudaMalloc(vid)");
    if ((e = cudaMalloc(&c->d_dead, std::max<size_t>(1, rows_cap))) != cudaSuccess) return fail(e, "cudaMalloc(dead)");
    if ((e = cudaMalloc(&c->d_tiles, std::max<size_t>(1, c->tiles.size()) * sizeof(TileDesc))) != cudaSuccess)
        return fail(e, "cudaMalloc(tiles)");
    if ((e = cudaMemset(c->d_dead, 0, std::max<size_t>(1, rows_cap))) != cudaSuccess) return fail(e, "cudaMemset(dead)");
    if ((e = cudaMemcpy(c->d_fp, fp.data(), fp.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(fp)");
    if ((e = cudaMemcpy(c->d_rec, rec.data(), rec.size() * sizeof(VerifyRec), cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(rec)");
    if (!ts.empty() && (e = cudaMemcpy(c->d_ts, ts.data(), ts.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(ts)");
    if ((e = cudaMemcpy(c->d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(off)");
    if (n_rows && (e = cudaMemcpy(c->d_vid, h_video_id, n_rows * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(vid)");
    if (!c->tiles.empty() &&
        (e = cudaMemcpy(c->d_tiles, c->tiles.data(), c->tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(tiles)");
    if (is_mutable) {
        // the tail starts as padding: fingerprint 0 may pass a filter, the NaN record behind it never verifies
        std::vector<VerifyRec> pad(static_cast<size_t>(c->tail_cap_vals), VerifyRec{kPadPattern, 0u, 0u});
        if ((e = cudaMemset(c->d_fp + c->n_units_main * kFpPerUnit, 0, c->tail_cap_vals * 2)) != cudaSuccess)
            return fail(e, "cudaMemset(tail fp)");
        if ((e = cudaMemcpy(c->d_rec + c->n_units_main * kFpPerUnit, pad.data(), pad.size() * sizeof(VerifyRec),
                            cudaMemcpyHostToDevice)) != cudaSuccess)
            return fail(e, "cudaMemcpy(tail rec)");
        if ((e = cudaStreamCreateWithFlags(&c->mut_stream, cudaStreamNonBlocking)) != cudaSuccess)
            return fail(e, "cudaStreamCreate");
        if ((e = cudaEventCreateWithFlags(&c->mut_event, cudaEventDisableTiming)) != cudaSuccess)
            return fail(e, "cudaEventCreate");
        c->tail_vals.reserve(static_cast<size_t>(c->tail_cap_vals));
        // the row an upsert replaces is found by video id: index the packed rows now (first row wins,
        // db.py:47 .first()), not inside the first upload's first add_timestamps()
        c->row_of_vid.reserve(static_cast<size_t>(n_rows) * 2 + 1024);
        for (long long r = 0; r < n_rows; ++r) c->row_of_vid.emplace(h_video_id[r], r);
        c->row_map_built = true;
    }
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail(e, "cudaDeviceSynchronize");
    *out = c;
    return TVZ_OK;
}

// Rebuild the tail from its live rows (host mirror) when it has filled up with replaced rows.
// Called with the lock held and no query in flight (the caller's contract for upserts).
int compact_tail(tvz_catalog *c) {
    std::vector<TailRow> rows;
    std::vector<unsigned long long> vals;
    for (const TailRow &r : c->tail_rows) {
        if (!r.alive) continue;
        TailRow n = r;
        n.start = static_cast<long long>(vals.size());
        vals.insert(vals.end(), c->tail_vals.begin() + r.start, c->tail_vals.begin() + r.start + r.n);
        rows.push_back(n);
    }
    const long long nv = static_cast<long long>(vals.size());
    const size_t cap = static_cast<size_t>(c->tail_cap_vals);
    std::vector<unsigned short> fp(cap, 0);
    std::vector<VerifyRec> rec(cap, VerifyRec{kPadPattern, 0u, 0u});
    std::vector<unsigned long long> ts(cap, kPadPattern);
    std::vector<long long> off(kTailRows + 1, c->ts_main_padded);
    std::vector<int> vid(kTailRows, 0);
    std::vector<unsigned char> dead(kTailRows, 0);
    for (size_t k = 0; k < rows.size(); ++k) {
        const long long row = c->n_rows_main + static_cast<long long>(k);
        for (int i = 0; i < rows[k].n; ++i) {
            const unsigned long long v = vals[rows[k].start + i];
            fp[rows[k].start + i] = static_cast<unsigned short>(filter_hash(v));
            rec[rows[k].start + i] = VerifyRec{v, static_cast<unsigned>(row), 0u};
            ts[rows[k].start + i] = v;
        }
        off[k] = c->ts_main_padded + rows[k].start;
        off[k + 1] = c->ts_main_padded + rows[k].start + rows[k].n;
        vid[k] = rows[k].vid;
        c->row_of_vid[rows[k].vid] = row;
    }
    cudaStream_t st = c->mut_stream;
    const long long p0 = c->n_units_main * kFpPerUnit;
    TVZ_CUDA(cudaMemcpyAsync(c->d_fp + p0, fp.data(), cap * 2, cudaMemcpyHostToDevice, st));
    TVZ_CUDA(cudaMemcpyAsync(c->d_rec + p0, rec.data(), cap * sizeof(VerifyRec), cudaMemcpyHostToDevice, st));
    TVZ_CUDA(cudaMemcpyAsync(c->d_ts + c->ts_main_padded, ts.data(), cap * 8, cudaMemcpyHostToDevice, st));
    TVZ_CUDA(cudaMemcpyAsync(c->d_off + c->n_rows_main, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
    TVZ_CUDA(cudaMemcpyAsync(c->d_vid + c->n_rows_main, vid.data(), vid.size() * 4, cudaMemcpyHostToDevice, st));
    TVZ_CUDA(cudaMemcpyAsync(c->d_dead + c->n_rows_main, dead.data(), dead.size(), cudaMemcpyHostToDevice, st));
    TVZ_CUDA(cudaStreamSynchronize(st));   // the host vectors go out of scope
    c->tail_rows.swap(rows);
    c->tail_vals.swap(vals);
    c->tail_used_vals = nv;
    c->tail_used_rows = static_cast<int>(c->tail_rows.size());
    return TVZ_OK;
}

}  // namespace

extern "C" {

int tvz_catalog_create(const double *h_ts, const int64_t *h_off, const int32_t *h_video_id, int64_t n_rows,
                       tvz_catalog **out) {
    return guarded([&]() -> int { return catalog_create(h_ts, h_off, h_video_id, n_rows, 0, out); });
}

int tvz_catalog_create_mutable(const double *h_ts, const int64_t *h_off, const int32_t *h_video_id, int64_t n_rows,
                               int64_t tail_values, tvz_catalog **out) {
    return guarded([&]() -> int {
        TVZ_REQUIRE(tail_values >= 1, "a mutable catalogue needs room for appended values");
        return catalog_create(h_ts, h_off, h_video_id, n_rows, tail_values, out);
    });
}

void tvz_catalog_destroy(tvz_catalog *c) {
    if (!c) return;
    if (c->mut_stream) cudaStreamSynchronize(c->mut_stream);
    free_catalog_device(c);
    delete c;
}

int64_t tvz_catalog_rows(const tvz_catalog *c) {
    if (!c) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    return c->n_rows();
}
int64_t tvz_catalog_values(const tvz_catalog *c) {
    if (!c) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    return c->n_vals_main + c->tail_used_vals;
}
int64_t tvz_catalog_algo_bytes(const tvz_catalog *c) {
    return c ? 8 * tvz_catalog_values(c) + 8 * (tvz_catalog_rows(c) + 1) : 0;
}
int tvz_catalog_tiles(const tvz_catalog *c) {
    if (!c) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    return static_cast<int>(c->n_tiles());
}

/* Row upsert (db.py:43-64): the first live row of `video_id` is replaced, or a new row is appended.
 * Contract: no query on this catalogue is in flight while an upsert runs (the Python Inspector holds a
 * reader/writer lock); queries enqueued afterwards see it.  TVZ_ERR_OVERFLOW: the tail is full even
 * after dropping replaced rows -- repack the catalogue. */
int tvz_catalog_upsert(tvz_catalog *c, int32_t video_id, const double *h_ts, int n) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(c && n >= 0 && (n == 0 || h_ts), "bad arguments");
    TVZ_REQUIRE(c->tail_index >= 0, "this catalogue was created immutable (tvz_catalog_create_mutable)");
    DeviceGuard on_device(c->device);
    std::vector<unsigned long long> vals;
    std::unordered_set<unsigned long long> seen;
    canon_row(h_ts, n, vals, seen);
    const int nv = static_cast<int>(vals.size());
    std::lock_guard<std::mutex> lk(c->mu);
    if (nv > c->tail_cap_vals)
        return set_error(TVZ_ERR_OVERFLOW, "a row of %d values does not fit the tail (%lld): repack", nv, c->tail_cap_vals);
    if (!c->row_map_built) {   // first upsert: index the packed rows by video id (first row wins, db.py:47)
        std::vector<int> vid(static_cast<size_t>(c->n_rows_main));
        if (c->n_rows_main) TVZ_CUDA(cudaMemcpy(vid.data(), c->d_vid, vid.size() * 4, cudaMemcpyDeviceToHost));
        c->row_of_vid.reserve(vid.size() * 2);
        for (long long r = 0; r < c->n_rows_main; ++r) c->row_of_vid.emplace(vid[r], r);
        c->row_map_built = true;
    }
    UpsertArgs u{};
    u.kill_row = -1;
    long long old_row = -1;
    auto it = c->row_of_vid.find(video_id);
    if (it != c->row_of_vid.end()) old_row = it->second;
    const bool old_is_last_tail = old_row >= c->n_rows_main && old_row == c->n_rows() - 1;
    // room: a replaced last tail row is rewritten in place, everything else is appended
    long long start = old_is_last_tail ? c->tail_rows.back().start : c->tail_used_vals;
    if (start + nv > c->tail_cap_vals || (!old_is_last_tail && c->tail_used_rows >= kTailRows)) {
        if (old_row >= c->n_rows_main) {   // it is being replaced anyway: do not carry it over
            c->tail_rows[old_row - c->n_rows_main].alive = false;
            c->row_of_vid.erase(video_id);
            old_row = -1;
        }
        int rc = compact_tail(c);
        if (rc) return rc;
        if (c->tail_used_vals + nv > c->tail_cap_vals || c->tail_used_rows >= kTailRows)
            return set_error(TVZ_ERR_OVERFLOW, "catalogue tail is full (%d live rows, %lld values): repack",
                             c->tail_used_rows, c->tail_used_vals);
        start = c->tail_used_vals;
        it = c->row_of_vid.find(video_id);
        old_row = it != c->row_of_vid.end() ? it->second : -1;
    }
    const bool in_place = old_row >= c->n_rows_main && old_row == c->n_rows() - 1;
    if (old_row >= 0) {
        u.kill_row = old_row;
        if (old_row < c->n_rows_main) {
            const long long a = c->h_off[old_row], b = c->h_off[old_row + 1];
            u.kill_pos_lo = a / kFpPerUnit * kFpPerUnit;
            u.kill_pos_hi = b > a ? (b + kFpPerUnit - 1) / kFpPerUnit * kFpPerUnit : u.kill_pos_lo;
            u.kill_ts_lo = a;
            u.kill_ts_hi = b;
            ++c->dead_main;
        } else {
            TailRow &tr = c->tail_rows[old_row - c->n_rows_main];
            u.kill_pos_lo = c->n_units_main * kFpPerUnit + tr.start;
            u.kill_pos_hi = u.kill_pos_lo + tr.n;
            u.kill_ts_lo = c->ts_main_padded + tr.start;
            u.kill_ts_hi = u.kill_ts_lo + tr.n;
            tr.alive = false;
        }
    }
    long long new_row;
    if (in_place) {
        new_row = old_row;
        TailRow &tr = c->tail_rows.back();
        tr.n = nv;
        tr.alive = true;
        c->tail_vals.resize(static_cast<size_t>(tr.start));
    } else {
        new_row = c->n_rows();
        c->tail_rows.push_back(TailRow{video_id, start, nv, true});
        ++c->tail_used_rows;
    }
    c->tail_vals.insert(c->tail_vals.end(), vals.begin(), vals.end());
    c->tail_used_vals = start + nv;
    c->row_of_vid[video_id] = new_row;
    u.rec = c->d_rec;
    u.fp = c->d_fp;
    u.ts = c->d_ts;
    u.off = c->d_off;
    u.vid = c->d_vid;
    u.dead = c->d_dead;
    u.new_row = new_row;
    u.pos0 = c->n_units_main * kFpPerUnit + start;
    u.ts0 = c->ts_main_padded + start;
    u.new_vid = video_id;
    u.n = nv;
    u.n_pos = c->n_pos_all;
    u.n_ts = c->n_ts_all;
    u.n_rows_cap = c->rows_cap;
    if (nv <= kParamKeys) {
        memcpy(u.vals, vals.data(), sizeof(unsigned long long) * nv);
    } else {
        if (nv > c->stage_cap) {
            TVZ_CUDA(cudaStreamSynchronize(c->mut_stream));
            if (c->d_stage) cudaFree(c->d_stage);
            if (c->h_stage) cudaFreeHost(c->h_stage);
            c->d_stage = nullptr;
            c->h_stage = nullptr;
            c->stage_cap = 0;
            long long cap = 1024;
            while (cap < nv) cap *= 2;
            TVZ_CUDA(cudaMalloc(&c->d_stage, cap * 8));
            TVZ_CUDA(cudaHostAlloc(&c->h_stage, cap * 8, cudaHostAllocDefault));
            c->stage_cap = cap;
        }
        TVZ_CUDA(cudaStreamSynchronize(c->mut_stream));   // the previous long row has left the staging buffer
        memcpy(c->h_stage, vals.data(), sizeof(unsigned long long) * nv);
        TVZ_CUDA(cudaMemcpyAsync(c->d_stage, c->h_stage, sizeof(unsigned long long) * nv, cudaMemcpyHostToDevice, c->mut_stream));
        u.vals_g = c->d_stage;
    }
    upsert_kernel<<<1, 512, 0, c->mut_stream>>>(u);
    TVZ_CUDA(cudaGetLastError());
    TVZ_CUDA(cudaEventRecord(c->mut_event, c->mut_stream));
    c->mut_seq.fetch_add(1, std::memory_order_relaxed);
    return TVZ_OK;
    });
}

/* {rows in the tail, values in the tail, value capacity, replaced rows of the packed part} */
int tvz_catalog_tail_info(const tvz_catalog *c, int64_t *out4) {
    TVZ_REQUIRE(c && out4, "null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    out4[0] = c->tail_used_rows;
    out4[1] = c->tail_used_vals;
    out4[2] = c->tail_cap_vals;
    out4[3] = c->dead_main;
    return TVZ_OK;
}

int tvz_match_ws_create(const tvz_catalog *cat, int64_t hit_capacity, tvz_match_ws **out) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(cat && out, "null pointer");
    *out = nullptr;
    TVZ_REQUIRE(hit_capacity >= 0, "negative capacity");
    tvz_match_ws *ws = new tvz_match_ws();
    ws->cat = cat;
    ws->cap = std::max<long long>(1, hit_capacity);
    auto bail = [&](cudaError_t e, const char *what) {
        set_error(TVZ_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        tvz_match_ws_destroy(ws);
        return TVZ_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    cudaStream_t st = ws->stream;
    const size_t n_state = static_cast<size_t>(std::max<long long>(1, cat->max_tiles())) * kBatch;
    if ((e = cudaMalloc(&ws->d_state, n_state * 4)) != cudaSuccess) return bail(e, "cudaMalloc(state)");
    if ((e = cudaMemsetAsync(ws->d_state, 0, n_state * 4, st)) != cudaSuccess) return bail(e, "cudaMemset(state)");
    if ((e = cudaMalloc(&ws->d_out, (ws->cap + 1) * 8)) != cudaSuccess) return bail(e, "cudaMalloc(out)");
    if ((e = cudaMemsetAsync(ws->d_out, 0, 8, st)) != cudaSuccess) return bail(e, "cudaMemset(out)");
    if ((e = cudaMalloc(&ws->d_rows, ws->cap * 8)) != cudaSuccess) return bail(e, "cudaMalloc(rows)");
    if ((e = cudaMalloc(&ws->d_kth, ws->cap * 4)) != cudaSuccess) return bail(e, "cudaMalloc(kth)");
    if ((e = cudaMalloc(&ws->d_nhits, 8 * kBatch)) != cudaSuccess) return bail(e, "cudaMalloc(nhits)");
    if ((e = cudaMemsetAsync(ws->d_nhits, 0, 8 * kBatch, st)) != cudaSuccess) return bail(e, "cudaMemset(nhits)");
    if ((e = cudaHostAlloc(&ws->h_out, (ws->cap + 1) * 8, cudaHostAllocDefault)) != cudaSuccess)
        return bail(e, "cudaHostAlloc(out)");
    if ((e = cudaHostAlloc(&ws->h_kth, ws->cap * 4, cudaHostAllocDefault)) != cudaSuccess)
        return bail(e, "cudaHostAlloc(kth)");
    if ((e = cudaEventCreateWithFlags(&ws->staged, cudaEventDisableTiming)) != cudaSuccess)
        return bail(e, "cudaEventCreate");
    int rc = ensure_query_capacity(ws, 256);
    if (rc) { tvz_match_ws_destroy(ws); return rc; }
    // queries may run on other streams: the initialisation must have landed before the first one
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return bail(e, "cudaStreamSynchronize");
    *out = ws;
    return TVZ_OK;
    });
}

void tvz_match_ws_destroy(tvz_match_ws *ws) {
    if (!ws) return;
    if (ws->stream) cudaStreamSynchronize(ws->stream);
    void *dev[] = {ws->d_state, ws->d_out, ws->d_rows, ws->d_kth, ws->d_nhits, ws->d_qcanon, ws->d_mult,
                   ws->b_out, ws->b_dev};
    for (void *p : dev)
        if (p) cudaFree(p);
    void *host[] = {ws->h_stage, ws->h_out, ws->h_kth, ws->hb_stage, ws->hb_out};
    for (void *p : host)
        if (p) cudaFreeHost(p);
    if (ws->staged) cudaEventDestroy(ws->staged);
    for (cudaEvent_t e : ws->b_staged)
        if (e) cudaEventDestroy(e);
    if (ws->t0) cudaEventDestroy(ws->t0);
    if (ws->t1) cudaEventDestroy(ws->t1);
    if (ws->stream) cudaStreamDestroy(ws->stream);
    delete ws;
}

const int32_t *tvz_match_ws_hits(const tvz_match_ws *ws) { return ws ? ws->d_out : nullptr; }
const int64_t *tvz_match_ws_nhits(const tvz_match_ws *ws) {
    return ws ? reinterpret_cast<const int64_t *>(ws->d_nhits) : nullptr;
}

// Debug hooks (not in the public header): time the kernel of the last query with CUDA
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
int tvz_debug_tile_trace(tvz_match_ws *ws, long long *d_trace) {   // int64 [tiles][16] on the device, or NULL to stop
    TVZ_REQUIRE(ws, "null workspace");
    ws->d_trace = d_trace;
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

// Debug hook (host only): the tiling of a catalogue with the given CSR offsets for `want` tiles.
// tiles_out: int32 [max_tiles][4] = {row_lo, n_rows, unit_lo, unit_hi}; returns the number of tiles.
int tvz_debug_build_tiles(const int64_t *off, int64_t n_rows, int want, int32_t *tiles_out, int max_tiles,
                          int32_t *uniform_units_out) {
    std::vector<long long> o(off, off + n_rows + 1);
    std::vector<TileDesc> t;
    const int U = build_tiles(o, n_rows, std::max(1, want), t);
    if (uniform_units_out) *uniform_units_out = U;
    for (size_t i = 0; i < t.size() && static_cast<int>(i) < max_tiles; ++i) {
        tiles_out[4 * i] = t[i].row_lo;
        tiles_out[4 * i + 1] = t[i].n_rows;
        tiles_out[4 * i + 2] = static_cast<int>(t[i].unit_lo);
        tiles_out[4 * i + 3] = static_cast<int>(t[i].unit_hi);
    }
    return static_cast<int>(t.size());
}

int tvz_catalog_match_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *h_q, int qn, int min_match,
                            int32_t *d_out, int64_t out_cap, void *stream) {
    return guarded([&]() -> int {
    return enqueue_match(cat, ws, h_q, qn, min_match, false, d_out, out_cap, static_cast<cudaStream_t>(stream));
    });
}

int tvz_catalog_match_gather_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *h_q, int qn,
                                   int min_match, int n_peers, int n_dst, const uint64_t *peer_record,
                                   const int32_t *d_my_slots, int64_t slot_stride_ints, int64_t out_cap, uint32_t epoch,
                                   void *stream) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(cat && cat->max_tiles() > 0, "the fused gather needs a non-empty shard");
    TVZ_REQUIRE(slot_stride_ints >= 4 * (out_cap + 1), "a gather slot holds 4 * (out_cap + 1) ints");
    GatherTargets gt;
    int rc = make_gather(n_peers, n_dst, peer_record, d_my_slots, slot_stride_ints, epoch, gt);
    if (rc) return rc;
    return enqueue_match(cat, ws, h_q, qn, min_match, false, nullptr, out_cap, static_cast<cudaStream_t>(stream), &gt);
    });
}

int tvz_catalog_match(const tvz_catalog *cat, tvz_match_ws *ws, const double *q, int qn, int min_match,
                      int32_t *out_video_id, int32_t *out_count, int32_t *out_kth, int64_t cap, int64_t *n_out) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(n_out, "null n_out");
    *n_out = 0;
    TVZ_REQUIRE(cap >= 0 && (cap == 0 || (out_video_id && out_count)), "bad output buffers");
    TVZ_REQUIRE(cat, "null catalogue");
    DeviceGuard on_device(cat->device);
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

/* Strided device -> host copy of a slice of n fixed-size records: bytes [offset, offset + width) of every
 * record, pitch_bytes apart on the device AND in h (same layout); sync != 0 waits for the stream.  The
 * sharded matchers read every shard's header plus an optimistic first slice of its hits this way. */
int tvz_copy_records_to_host(const void *d_rec, void *h_rec, int n_records, int64_t pitch_bytes, int64_t offset_bytes,
                             int64_t width_bytes, int sync, void *stream) {
    TVZ_REQUIRE(d_rec && h_rec && n_records >= 1 && offset_bytes >= 0 && width_bytes >= 0 &&
                offset_bytes + width_bytes <= pitch_bytes, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (width_bytes)
        TVZ_CUDA(cudaMemcpy2DAsync(static_cast<char *>(h_rec) + offset_bytes, pitch_bytes,
                                   static_cast<const char *>(d_rec) + offset_bytes, pitch_bytes, width_bytes, n_records,
                                   cudaMemcpyDeviceToHost, st));
    if (sync) TVZ_CUDA(cudaStreamSynchronize(st));
    return TVZ_OK;
}

int tvz_catalog_batch_limit(void) { return kParamKeys; }
int tvz_catalog_batch_size(void) { return kBatch; }

/* Device-resident batch: up to 8 queries, records to d_out int32 [n_queries][out_cap + 1][2]. */
int tvz_catalog_match_batch_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *q_all, const int64_t *q_off,
                                  int n_queries, int min_match, int32_t *d_out, int64_t out_cap, void *stream) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(cat && ws && ws->cat == cat, "workspace does not belong to this catalogue");
    TVZ_REQUIRE(q_off && d_out && (q_all || q_off[n_queries] == 0), "bad arguments");
    return enqueue_batch(cat, ws, q_all, q_off, 0, n_queries, min_match, d_out, out_cap, static_cast<cudaStream_t>(stream));
    });
}

int tvz_catalog_match_batch_gather_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *q_all,
                                         const int64_t *q_off, int n_queries, int min_match, int n_peers, int n_dst,
                                         const uint64_t *peer_record, const int32_t *d_my_slots,
                                         int64_t slot_stride_ints, int64_t out_cap, uint32_t epoch, void *stream) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(cat && ws && ws->cat == cat, "workspace does not belong to this catalogue");
    TVZ_REQUIRE(cat->max_tiles() > 0, "the fused gather needs a non-empty shard");
    TVZ_REQUIRE(q_off && (q_all || q_off[n_queries] == 0), "bad arguments");
    TVZ_REQUIRE(slot_stride_ints >= static_cast<int64_t>(kBatch) * 4 * (out_cap + 1), "a batched gather slot holds 8 * 4 * (out_cap + 1) ints");
    GatherTargets gt;
    int rc = make_gather(n_peers, n_dst, peer_record, d_my_slots, slot_stride_ints, epoch, gt);
    if (rc) return rc;
    return enqueue_batch(cat, ws, q_all, q_off, 0, n_queries, min_match, nullptr, out_cap, static_cast<cudaStream_t>(stream),
                         &gt);
    });
}

/* Any number of queries, 8 per catalogue pass; host buffers in and out.  Queries with more than
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
    DeviceGuard on_device(cat->device);
    cudaStream_t st = ws->stream;
    const long long rec = (ws->cap + 1) * 2;
    long long written = 0;
    bool overflow = false;
    long long need_cap = 0, need_total = 0;
    for (int g0 = 0; g0 < n_queries; g0 += kBatch) {
        const int nb = std::min(kBatch, n_queries - g0);
        int rc = enqueue_batch(cat, ws, q_all, q_off, g0, nb, min_match, nullptr, 0, st);
        if (rc) return rc;
        // headers + an optimistic first slice of every query's hits in one strided copy
        const long long first = std::min<long long>(ws->cap, 512);
        TVZ_CUDA(cudaMemcpy2DAsync(ws->hb_out, rec * 4, ws->b_out, rec * 4, (first + 1) * 8, nb, cudaMemcpyDeviceToHost, st));
        TVZ_CUDA(cudaStreamSynchronize(st));
        for (bool &busy : ws->b_stage_busy) busy = false;
        bool more = false;
        for (int b = 0; b < nb; ++b) {
            const long long n = ws->hb_out[b * rec];
            if (n > ws->cap) { overflow = true; need_cap = std::max(need_cap, n); continue; }
            if (n > first) {
                TVZ_CUDA(cudaMemcpyAsync(ws->hb_out + b * rec + 2 * (first + 1), ws->b_out + b * rec + 2 * (first + 1),
                                         (n - first) * 8, cudaMemcpyDeviceToHost, st));
                more = true;
            }
        }
        if (more) TVZ_CUDA(cudaStreamSynchronize(st));
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
