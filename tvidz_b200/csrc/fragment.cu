// Fragment (offset-invariant) matching -- SURVEY.md section 8 row a8 / BASELINE config 5.
//
// The reference advertises fragment detection (README.md:5) but implements none
// (inspector/db.py:78-79 is exact membership at offset 0), so the semantics are defined HERE
// and checked only against this repo's own CPU restatement (test infrastructure): PARITY UNPINNED.
//
// Spec ("interval-anchored alignment"), on integer ticks t = llround(ts * tick_hz):
//   * rows and the query are sets of ticks, sorted ascending, duplicates removed;
//   * a candidate offset d = C[j] - Q[i] is generated wherever `anchor` CONSECUTIVE intervals
//     agree:  | (C[j+a+1]-C[j+a]) - (Q[i+a+1]-Q[i+a]) | <= tol_gap  for a = 0 .. anchor-1
//     (anchor = 0: EVERY pair (i, j) is a candidate -- SURVEY.md B.4 exactly as written, the
//     exhaustive mode the anchored ones are measured against (DESIGN.md recall table);
//     anchor = 1: every agreeing pair of adjacent cuts; anchor = 2, the default of the Python API:
//     two agreeing intervals in a row, i.e. three cuts);
//   * score(d) = #{ i : exists j with | Q[i] + d - C[j] | <= tol };
//   * the row's verdict is the candidate with the highest score (ties: smaller |d|, then
//     smaller d); the row is reported iff that score >= min_match, as
//     (video_id, score, d);  offset in seconds = d / tick_hz.
//   * zero_offset_only = 1 replaces the candidate set by {0}: with tol = 0 the score is
//     then the reference's match_count (db.py:85-89) on tick-exact data.
//
// Two kernels.  anchor >= 2 (fragment_stream_kernel, below): agreeing interval n-grams are rare, so
// the whole catalogue is STREAMED as one flat int32 array -- 256-bit loads, one bucket-table lookup
// per stored tick, rows resolved only for the few survivors -- and the kernel is HBM-bound.
// anchor = 1 / zero_offset_only (fragment_match_kernel): one warp per row.  The row's int32 ticks are staged into a per-warp shared-memory
// buffer with coalesced loads (4 B per stored timestamp is all the HBM traffic).  Every row
// interval is tested against a 4 KB bitmap of the (tolerance-widened) query intervals; the
// ~10% that survive look up their matching query intervals in the sorted interval list, the
// resulting candidates (offset + anchor) are parked in shared memory through a warp prefix sum
// and scored 32 at a time by a merge walk outwards from the anchor; a warp shuffle reduction
// with the spec's tie-breaks picks the row's best.
// Row results land in score[row] / delta[row]; the ordered compaction of match.cu turns them
// into the hit list.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace tvz {
namespace {

constexpr int kFragWarps = 8;
constexpr int kFragThreads = kFragWarps * 32;
constexpr int kCandCap = 128;       // candidate offsets parked per warp
constexpr int kFragMaxQ = 256;      // query cuts per launch (1 KB of kernel parameters)
constexpr int kGapRange = 1 << 15;  // intervals up to 32767 ticks get their own bucket; longer ones share the last
constexpr int kMaxBuckets = 2048;   // interval buckets (bucket width = power of two > tol_gap, at least 16 ticks)
constexpr int kSpan = 64;           // row intervals per lane per pass (one 64-bit hit mask)

struct FragQuery {
    int qn, tol, tol_gap, zero_only, min_match, exhaustive;
    int q[kFragMaxQ];
};

struct FragSmem {
    int cand_d[kFragWarps][kCandCap];
    int cand_ij[kFragWarps][kCandCap];       // i << 24 | j
    int n_cand[kFragWarps];
    unsigned short bucket[kMaxBuckets];      // start << 8 | count: sorted query intervals that may match
    int q[kFragMaxQ];
    int sgap[kFragMaxQ];                     // query intervals, ascending
    int sidx[kFragMaxQ];                     // ... and the query position each came from
};

struct Best {
    int score, delta;
};
__device__ __forceinline__ bool better(int s, int d, const Best &b) {
    if (s != b.score) return s > b.score;
    const int ad = d < 0 ? -d : d, ab = b.delta < 0 ? -b.delta : b.delta;
    if (ad != ab) return ad < ab;
    return d < b.delta;
}

// Ticks are bounded by 5e8 in magnitude (to_tick), so Q[i] + d +- tol stays inside int32.

// score(d) by binary search + merge walk (zero-offset mode).
__device__ __forceinline__ int score_offset(const int *__restrict__ C, int L, const int *__restrict__ Q, int qn, int d,
                                            int tol) {
    const int first = Q[0] + d - tol;
    int lo = 0, hi = L;  // first C[p] >= first
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(C + mid) < first) lo = mid + 1; else hi = mid;
    }
    int p = lo, s = 0;
    for (int i = 0; i < qn && p < L; ++i) {
        const int t = Q[i] + d;
        while (p < L && __ldg(C + p) < t - tol) ++p;
        if (p < L && __ldg(C + p) <= t + tol) ++s;
    }
    return s;
}

// score(d) for a candidate anchored at Q[i] + d == C[j]: merge walks forwards and backwards from
// the anchor.  Each walk is ONE flat loop whose iteration either steps the row pointer or the query
// cut, so the 32 candidates a warp scores together run nearly the same number of iterations.
__device__ __forceinline__ int score_anchored(const int *__restrict__ C, int L, const int *__restrict__ Q, int qn, int d,
                                              int tol, int i, int j, int need) {
    // `need`: a candidate that cannot reach this score is irrelevant (it can neither beat the lane's
    // best nor be reported); its walk is abandoned as soon as s + (cuts not yet examined) < need.
    // Each loop iteration is branch-free: it either steps the row pointer (row cut still below the
    // window of query cut k) or settles cut k and moves to the next one.
    int s = 1, left = qn - 1;
    const int span = 2 * tol;
    {
        int p = j, k = i + 1;           // start AT the anchor: cuts closer than tol may share a row cut
        while (k < qn && p < L && s + left >= need) {
            const int lo = Q[k] + d - tol;          // window of query cut k: [lo, lo + 2 tol]
            const int c = __ldg(C + p);
            const int below = c < lo;
            s += (1 - below) & (c <= lo + span);    // first C[p] >= lo decides cut k; C[p] may also serve cut k+1
            left -= 1 - below;
            k += 1 - below;
            p += below;
        }
        left -= qn - k;                 // cuts past the end of the row cannot match
    }
    {
        int p = j, k = i - 1;
        while (k >= 0 && p >= 0 && s + left >= need) {
            const int hi = Q[k] + d + tol;
            const int c = __ldg(C + p);
            const int above = c > hi;
            s += (1 - above) & (c >= hi - span);
            left -= 1 - above;
            k -= 1 - above;
            p -= above;
        }
    }
    return s;
}

__global__ void __launch_bounds__(kFragThreads)
fragment_match_kernel(const int *__restrict__ ticks, const long long *__restrict__ off, long long n_rows,
                      const __grid_constant__ FragQuery fq, int shift, int *__restrict__ score_out,
                      int *__restrict__ delta_out) {
    __shared__ FragSmem sm;
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qn = fq.qn, tol = fq.tol, tol_gap = fq.tol_gap;
    const int ng = qn > 0 ? qn - 1 : 0;  // query intervals
    const int n_buckets = kGapRange >> shift;
    for (int i = threadIdx.x; i < qn; i += kFragThreads) sm.q[i] = fq.q[i];
    if (threadIdx.x < kFragWarps) sm.n_cand[threadIdx.x] = 0;
    __syncthreads();
    // query intervals sorted ascending by rank counting (ng <= 255)
    for (int i = threadIdx.x; i < ng; i += kFragThreads) {
        const int g = sm.q[i + 1] - sm.q[i];
        int rank = 0;
        for (int k = 0; k < ng; ++k) {
            const int gk = sm.q[k + 1] - sm.q[k];
            rank += (gk < g) || (gk == g && k < i);
        }
        sm.sgap[rank] = g;
        sm.sidx[rank] = i;
    }
    __syncthreads();
    // bucket B covers interval lengths [B << shift, (B + 1) << shift) (the last one: everything above);
    // its entry is the contiguous range of sorted query intervals within tol_gap of any length in it
    for (int B = threadIdx.x; B < n_buckets; B += kFragThreads) {
        const long long lo_len = (static_cast<long long>(B) << shift) - tol_gap;
        const long long hi_len = B == n_buckets - 1 ? (1ll << 40) : ((static_cast<long long>(B + 1) << shift) - 1 + tol_gap);
        int a = 0, z = ng;
        while (a < z) { const int m = (a + z) >> 1; if (sm.sgap[m] < lo_len) a = m + 1; else z = m; }
        const int first = a;
        z = ng;
        while (a < z) { const int m = (a + z) >> 1; if (sm.sgap[m] <= hi_len) a = m + 1; else z = m; }
        sm.bucket[B] = static_cast<unsigned short>((first << 8) | (a - first));
    }
    __syncthreads();
    int *cand_d = sm.cand_d[warp];
    int *cand_ij = sm.cand_ij[warp];
    int *n_cand = &sm.n_cand[warp];
    const long long warps = static_cast<long long>(gridDim.x) * kFragWarps;

    for (long long r = static_cast<long long>(blockIdx.x) * kFragWarps + warp; r < n_rows; r += warps) {
        const long long b = off[r];
        const int L = static_cast<int>(off[r + 1] - b);
        const int *C = ticks + b;  // read through L1: the row (a few KB) stays cached while it is worked on
        Best best{0, 0};
        auto verify = [&]() {  // parked candidates -> lane-local best
            __syncwarp();
            const int n = min(*n_cand, kCandCap);
            for (int c = lane; c < n; c += 32) {
                const int d = cand_d[c], ij = cand_ij[c];
                const int s = score_anchored(C, L, sm.q, qn, d, tol, static_cast<int>(static_cast<unsigned>(ij) >> 24),
                                             ij & 0xffffff, max(fq.min_match, best.score));
                if (better(s, d, best)) best = Best{s, d};
            }
            __syncwarp();
            if (lane == 0) *n_cand = 0;
            __syncwarp();
        };
        if (qn > 0 && L > 0) {
            if (fq.zero_only) {
                if (lane == 0) best = Best{score_offset(C, L, sm.q, qn, 0, tol), 0};
            } else if (fq.exhaustive) {
                // SURVEY.md B.4 as written: EVERY offset d = C[j] - Q[i] is a candidate.  A lane takes row
                // cuts j = lane, lane + 32, ...; a walk is abandoned (exactly) once it cannot reach
                // max(min_match, best so far).
                for (int j = lane; j < L; j += 32) {
                    const int c0 = __ldg(C + j);
                    for (int i = 0; i < qn; ++i) {
                        const int d = c0 - sm.q[i];
                        const int s = score_anchored(C, L, sm.q, qn, d, tol, i, j, max(fq.min_match, best.score));
                        if (better(s, d, best)) best = Best{s, d};
                    }
                }
            } else if (ng > 0) {
                for (int base = 0; base < L - 1; base += 32 * kSpan) {
                    // pass 1 (branch free): which of this lane's intervals fall in a non-empty bucket
                    unsigned long long hits = 0;
                    const int steps = min(kSpan, (L - 1 - base + 31) >> 5);
                    for (int t = 0; t < steps; ++t) {
                        const int j = base + t * 32 + lane;
                        const bool valid = j < L - 1;
                        const int g = valid ? __ldg(C + j + 1) - __ldg(C + j) : 0;
                        const unsigned e = sm.bucket[min(g >> shift, n_buckets - 1)];
                        hits |= static_cast<unsigned long long>(valid && (e & 0xffu)) << t;
                    }
                    // pass 2: every lane works through its own hits; candidates are parked per warp
                    while (__any_sync(0xffffffffu, hits != 0)) {
                        if (hits) {
                            const int t = __ffsll(static_cast<long long>(hits)) - 1;
                            hits &= hits - 1;
                            const int j = base + t * 32 + lane;
                            const int c0 = __ldg(C + j);
                            const int g = __ldg(C + j + 1) - c0;
                            const unsigned e = sm.bucket[min(g >> shift, n_buckets - 1)];
                            const int k0 = e >> 8, k1 = k0 + (e & 0xffu);
                            for (int k = k0; k < k1; ++k) {
                                const int diff = g - sm.sgap[k];
                                if (diff > tol_gap || diff < -tol_gap) continue;
                                const int i = sm.sidx[k], d = c0 - sm.q[i];
                                const int slot = atomicAdd(n_cand, 1);
                                if (slot < kCandCap) {
                                    cand_d[slot] = d;
                                    cand_ij[slot] = (i << 24) | j;
                                } else {  // parking lot full: score in place
                                    const int s = score_anchored(C, L, sm.q, qn, d, tol, i, j, max(fq.min_match, best.score));
                                    if (better(s, d, best)) best = Best{s, d};
                                }
                            }
                        }
                        __syncwarp();
                        if (*n_cand >= kCandCap - 32) verify();   // each lane adds at most a few per round
                    }
                }
                verify();
            }
        }
        // warp reduction of (score, delta) with the spec's tie-breaks
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int s = __shfl_xor_sync(0xffffffffu, best.score, o);
            const int d = __shfl_xor_sync(0xffffffffu, best.delta, o);
            if (better(s, d, best)) best = Best{s, d};
        }
        if (lane == 0) {
            score_out[r] = best.score;
            delta_out[r] = best.delta;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// anchor >= 2: streaming kernel.
//
// The tick array of the whole shard is read ONCE, front to back, regardless of row boundaries
// (4 B per stored timestamp; `off` is touched only for survivors).  A warp owns 1536 consecutive
// ticks per iteration as six 256-bit loads per lane (rolling prefetch: a unit's registers are
// reloaded from the next chunk as soon as it has been tested; 16 warps x 6 KB in flight per SM,
// 128 registers so that nothing spills -- measured sweep in profiles/r01_fragment_stream_sweep.txt).  Per tick: the interval to the next
// tick indexes a 2048-entry shared-memory table whose entry is a 32-bit set of the query intervals
// (index mod 32) that could agree with an interval in that bucket; position p survives iff some
// query interval i is in the set of p, i+1 in the set of p+1 (, i+2 in the set of p+2):
//     hit(p) = m[p] & rotr(m[p+1], 1) [& rotr(m[p+2], 2)]
// -- a rotate and an AND per tick, one warp vote per 256 ticks.  Survivors (~0.05 % of positions)
// are parked in a per-warp shared-memory queue and resolved when the warp has finished streaming
// (stream_drain).  Intervals that straddle two rows are garbage and are rejected with everything
// else there: row through the coarse index + a short search in `off`, exact interval test against
// every query position the set names, then the same anchored merge walk as the per-row kernel
// scores the offset and the row's best candidate is combined across warps with one 64-bit
// atomicMax (frag_key).
// launch shape (overridable for tuning sweeps: scripts/sweep_fragment.py builds variants)
#ifndef TVZ_FS_THREADS
#define TVZ_FS_THREADS 256
#endif
#ifndef TVZ_FS_MINB
#define TVZ_FS_MINB 2
#endif
#ifndef TVZ_FS_UNITS
#define TVZ_FS_UNITS 6
#endif
constexpr int kStreamThreads = TVZ_FS_THREADS;
constexpr int kStreamMinBlocks = TVZ_FS_MINB;
constexpr int kStreamWarps = kStreamThreads / 32;
constexpr int kStreamUnits = TVZ_FS_UNITS;                          // 256-bit loads in flight per thread
constexpr int kUnitTicks = 32 * 8;                       // one warp-wide 256-bit load
constexpr int kWarpChunk = kStreamUnits * kUnitTicks;    // 1536 consecutive ticks per warp per iteration
constexpr int kCtaChunk = kStreamWarps * kWarpChunk;     // 12288 ticks = 48 KB
constexpr int kStreamQueue = 256 + 64;                   // survivors parked per warp (>= the 256 one unit can add)
constexpr int kStreamBlockShift = 8;                     // coarse row index: one entry per 256 ticks
constexpr int kMaxAnchor = 3;
constexpr int kPadTick = 0x7fffffff;

struct StreamSmem {
    unsigned table[kMaxBuckets];                    // bucket -> set of query intervals (index mod 32)
    int q[kFragMaxQ];
    long long qpos[kStreamWarps][kStreamQueue];
    unsigned qset[kStreamWarps][kStreamQueue];
};

struct I32x8 {
    int t[8];
};
__device__ __forceinline__ I32x8 ld_stream_ticks(const int *p) {  // LDG.E.256, no L1 allocation, evict-first in L2
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
    I32x8 r;
    r.t[0] = static_cast<int>(a); r.t[1] = static_cast<int>(a >> 32);
    r.t[2] = static_cast<int>(b); r.t[3] = static_cast<int>(b >> 32);
    r.t[4] = static_cast<int>(c); r.t[5] = static_cast<int>(c >> 32);
    r.t[6] = static_cast<int>(d); r.t[7] = static_cast<int>(d >> 32);
    return r;
}
__device__ __forceinline__ unsigned rotr32(unsigned x, int n) { return __funnelshift_r(x, x, n); }

// Everything a survivor needs to be parked and resolved; lives in local memory and is touched only
// on the rare path.
struct StreamCtx {
    const int *ticks;
    const long long *off;
    const int *block_row;
    unsigned long long *keys;
    const int *q;          // shared memory
    long long *qpos;       // this warp's queue (shared memory)
    unsigned *qset;
    long long n_vals;
    int qn, tol, tol_gap, min_match;
};

// Survivor at flat position `pos` whose set names the query intervals (mod 32) that may start an
// agreeing run of A intervals there: find its row, test exactly, score, combine into keys[row].
template <int A>
__device__ __forceinline__ void stream_resolve(const StreamCtx &cx, long long pos, unsigned set) {
    if (pos + A >= cx.n_vals) return;  // padding
    const long long b = pos >> kStreamBlockShift;
    long long a = cx.block_row[b], z = cx.block_row[b + 1] + 1;  // last row with off[row] <= pos is in [a, z)
    while (z - a > 1) {
        const long long mid = (a + z) >> 1;
        if (cx.off[mid] <= pos) a = mid; else z = mid;
    }
    const long long rb = cx.off[a], re = cx.off[a + 1];
    if (pos + A >= re) return;          // the anchor's A + 1 cuts must all belong to this row
    const int *C = cx.ticks + rb;
    const int L = static_cast<int>(re - rb), j = static_cast<int>(pos - rb);
    int cg[A];
#pragma unroll
    for (int k = 0; k < A; ++k) cg[k] = __ldg(C + j + k + 1) - __ldg(C + j + k);
    const int c0 = __ldg(C + j);
    unsigned long long best = 0;
    while (set) {
        const int bit = __ffs(set) - 1;
        set &= set - 1;
        for (int i = bit; i + A < cx.qn; i += 32) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < A; ++k) {
                const int diff = cg[k] - (cx.q[i + k + 1] - cx.q[i + k]);
                ok = ok && diff <= cx.tol_gap && diff >= -cx.tol_gap;
            }
            if (!ok) continue;
            const int d = c0 - cx.q[i];
            const int s = score_anchored(C, L, cx.q, cx.qn, d, cx.tol, i, j, cx.min_match);
            if (s >= cx.min_match) best = max(best, frag_key(s, d));
        }
    }
    if (best) atomicMax(&cx.keys[a], best);
}

// Resolve the warp's parked survivors, 32 at a time (called by the whole warp).  Normally this
// runs ONCE, after the warp's last chunk: a warp meets a few dozen survivors over its whole share
// of the catalogue, the queue holds 320, and at that point every warp of the grid is resolving at
// the same time, so the dependent loads of one survivor hide behind those of ~75,000 others.
// Only a query that matches nearly everywhere fills the queue earlier.
template <int A>
__device__ __noinline__ void stream_drain(const StreamCtx &cx, int n) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    for (int i = lane; i < n; i += 32) stream_resolve<A>(cx, cx.qpos[i], cx.qset[i]);
    __syncwarp();
}

template <int A>
__global__ void __launch_bounds__(kStreamThreads, kStreamMinBlocks)
fragment_stream_kernel(const int *__restrict__ ticks, long long n_vals, long long n_padded,
                       const long long *__restrict__ off, const int *__restrict__ block_row, long long n_rows,
                       const __grid_constant__ FragQuery fq, int shift, unsigned long long *__restrict__ keys,
                       int l2_ahead, int queue_cap) {
    static_assert(A >= 2 && A <= kMaxAnchor, "anchor length");
    __shared__ StreamSmem sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qn = fq.qn;
    const int ng = qn - 1;                      // query intervals (>= A, checked by the host)
    for (int i = threadIdx.x; i < kMaxBuckets; i += kStreamThreads) sm.table[i] = 0u;
    for (int i = threadIdx.x; i < qn; i += kStreamThreads) sm.q[i] = fq.q[i];
    __syncthreads();
    // An interval of length g lives in bucket (g >> shift) mod 2048 -- long intervals (and the negative
    // garbage between two rows) simply wrap around; query interval i marks every bucket that holds a
    // length within tol_gap of it, so wrapping can add false survivors but never lose a true one.
    for (int i = threadIdx.x; i < ng; i += kStreamThreads) {
        const long long g = static_cast<long long>(sm.q[i + 1]) - sm.q[i];
        const long long lo = max(0ll, g - fq.tol_gap) >> shift, hi = (g + fq.tol_gap) >> shift;  // <= 3 buckets
        for (long long b = lo; b <= hi; ++b) atomicOr(&sm.table[b & (kMaxBuckets - 1)], 1u << (i & 31));
    }
    __syncthreads();
    pdl_wait();  // keys[] is being read and zeroed by the previous query's compaction until here
    pdl_launch_dependents();

    const StreamCtx cx{ticks, off, block_row, keys, sm.q, sm.qpos[warp], sm.qset[warp], n_vals, qn, fq.tol, fq.tol_gap,
                       fq.min_match};
    long long *qpos = sm.qpos[warp];
    unsigned *qset = sm.qset[warp];
    int queued = 0;  // warp-uniform, <= kStreamQueue

    // byte offset of the bucket straight from the interval: one shift, one mask, LDS [reg + imm]
    const int byte_shift = shift - 2;
    auto bucket_set = [&](int g) -> unsigned {
        const unsigned byte = (static_cast<unsigned>(g) >> byte_shift) & ((kMaxBuckets - 1) << 2);
        return *reinterpret_cast<const unsigned *>(reinterpret_cast<const unsigned char *>(sm.table) + byte);
    };

    const long long stride = static_cast<long long>(gridDim.x) * kCtaChunk;
    long long base = static_cast<long long>(blockIdx.x) * kCtaChunk + warp * kWarpChunk;  // this warp's first tick
    I32x8 v[kStreamUnits];
    if (base < n_padded) {
#pragma unroll
        for (int u = 0; u < kStreamUnits; ++u) v[u] = ld_stream_ticks(ticks + base + u * kUnitTicks + lane * 8);
    }
    const int next_lane = (lane + 1) & 31;
    int la[A], la_next[A];  // the A ticks behind this warp's chunk (the array is padded past n_padded)
#pragma unroll
    for (int k = 0; k < A; ++k) la_next[k] = base < n_padded ? __ldg(ticks + base + kWarpChunk + k) : 0;
    for (; base < n_padded; base += stride) {
        const bool more = base + stride < n_padded;
        // optional (off by default, it did not pay): HBM -> L2 `l2_ahead` iterations ahead, one TMA
        // prefetch of the warp's whole chunk
        if (l2_ahead > 0 && lane == 0 && base + l2_ahead * stride < n_padded)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ticks + base + l2_ahead * stride),
                         "r"(kWarpChunk * 4)
                         : "memory");
#pragma unroll
        for (int k = 0; k < A; ++k) {
            la[k] = la_next[k];
            if (more) la_next[k] = __ldg(ticks + base + stride + kWarpChunk + k);
        }
#pragma unroll
        for (int u = 0; u < kStreamUnits; ++u) {
            // Lane L needs tick 8 and the sets 8 .. 8+A-2 of its run from lane L+1; lane 31 from lane 0
            // of the NEXT unit, whose registers lane 0 still holds -- lane 0 publishes those instead of
            // its own (nobody reads lane 0's own).
            int nx[A];
#pragma unroll
            for (int k = 0; k < A; ++k) nx[k] = u + 1 < kStreamUnits ? v[(u + 1) % kStreamUnits].t[k] : la[k];
            unsigned m[8 + A - 1];
            const int t8 = __shfl_sync(0xffffffffu, lane == 0 ? nx[0] : v[u].t[0], next_lane);
#pragma unroll
            for (int k = 0; k < 7; ++k) m[k] = bucket_set(v[u].t[k + 1] - v[u].t[k]);
            m[7] = bucket_set(t8 - v[u].t[7]);
#pragma unroll
            for (int k = 0; k < A - 1; ++k) {
                const unsigned nxset = bucket_set(nx[k + 1] - nx[k]);
                m[8 + k] = __shfl_sync(0xffffffffu, lane == 0 ? nxset : m[k], next_lane);
            }
            unsigned acc = 0;  // OR of the 8 positions' surviving sets: a rotate and one LOP3 per tick
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                unsigned hit = m[k];
#pragma unroll
                for (int a2 = 1; a2 < A - 1; ++a2) hit &= rotr32(m[k + a2], a2);
                acc |= hit & rotr32(m[k + A - 1], A - 1);
            }
#ifdef TVZ_FS_NOPARK
            if (acc == 0xdeadbeefu) {
#else
            if (__any_sync(0xffffffffu, acc != 0u)) {
#endif
                // Rare path (inline, registers only): every lane re-derives which of its 8 positions
                // survived, a shuffle scan hands out queue slots, the lanes park (position, set).
                unsigned flags = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    unsigned hit = m[k];
#pragma unroll
                    for (int a2 = 1; a2 < A; ++a2) hit &= rotr32(m[k + a2], a2);
                    flags |= (hit != 0u) << k;
                }
                const int mine = __popc(flags);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += n;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (queued + total > queue_cap) {  // make room (a unit adds at most 256, the queue holds 320)
                    stream_drain<A>(cx, queued);
                    queued = 0;
                }
                int slot = queued + incl - mine;
                const long long pos = base + u * kUnitTicks + lane * 8;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (flags & (1u << k)) {
                        unsigned hit = m[k];
#pragma unroll
                        for (int a2 = 1; a2 < A; ++a2) hit &= rotr32(m[k + a2], a2);
                        qpos[slot] = pos + k;
                        qset[slot] = hit;
                        ++slot;
                    }
                }
                queued += total;
            }
            if (more) v[u] = ld_stream_ticks(ticks + base + stride + u * kUnitTicks + lane * 8);
        }
    }
    stream_drain<A>(cx, queued);
}

}  // namespace
}  // namespace tvz

using namespace tvz;

struct tvz_fragcat {
    int device = 0;
    double tick_hz = 1000.0;
    long long n_rows = 0, n_vals = 0;
    long long n_padded = 0;                       // ticks are padded to whole streaming chunks (+ 8 look-ahead)
    int *d_ticks = nullptr;
    long long *d_off = nullptr;
    int *d_vid = nullptr;
    int *d_block_row = nullptr;                   // row holding tick b * 256 (streaming kernel: survivor -> row)
    unsigned long long *d_keys = nullptr;         // [n_rows] best-candidate keys, zero between queries
    // workspace (calls on one catalogue are serialised by `mu`)
    std::mutex mu;
    long long cap = 0;
    int *d_score = nullptr, *d_delta = nullptr;   // [n_rows]
    int *d_out = nullptr;                         // [cap+1][2] records, then [cap+1] deltas
    long long *d_rows = nullptr, *d_nhits = nullptr;
    unsigned long long *d_state = nullptr;
    unsigned *d_ticket = nullptr;
    int *h_out = nullptr;                         // pinned mirror of d_out
    cudaStream_t stream = nullptr;
};

namespace {

// llround(ts * tick_hz) as int32, or false if not representable / not finite
inline bool to_tick(double ts, double hz, int *out) {
    const double v = ts * hz;
    if (!std::isfinite(v) || std::fabs(v) > 5.0e8) return false;  // Q + d +- tol stays inside int32
    *out = static_cast<int>(std::llround(v));
    return true;
}

int frag_reserve(tvz_fragcat *c, long long cap) {
    if (cap <= c->cap) return TVZ_OK;
    if (c->d_out) cudaFree(c->d_out);
    if (c->d_rows) cudaFree(c->d_rows);
    if (c->h_out) cudaFreeHost(c->h_out);
    c->d_out = nullptr; c->d_rows = nullptr; c->h_out = nullptr; c->cap = 0;
    TVZ_CUDA(cudaMalloc(&c->d_out, (cap + 1) * 12));
    TVZ_CUDA(cudaMemset(c->d_out, 0, 8));
    TVZ_CUDA(cudaMalloc(&c->d_rows, cap * 8));
    TVZ_CUDA(cudaHostAlloc(&c->h_out, (cap + 1) * 12, cudaHostAllocDefault));
    TVZ_CUDA(cudaDeviceSynchronize());
    c->cap = cap;
    return TVZ_OK;
}

int frag_enqueue(tvz_fragcat *c, const double *h_q, int qn, int min_match, int tol, int tol_gap, int anchor,
                 int zero_only, int *d_out, long long out_cap, cudaStream_t st, const GatherTargets *gather = nullptr) {
    TVZ_REQUIRE(qn >= 0 && (qn == 0 || h_q), "bad query");
    TVZ_REQUIRE(tol >= 0 && tol_gap >= 0, "negative tolerance");
    TVZ_REQUIRE(anchor >= 0 && anchor <= kMaxAnchor, "anchor_intervals must be 0 (exhaustive) .. %d", kMaxAnchor);
    FragQuery fq{};
    std::vector<int> q;
    q.reserve(qn);
    for (int i = 0; i < qn; ++i) {
        int t;
        if (to_tick(h_q[i], c->tick_hz, &t)) q.push_back(t);
    }
    std::sort(q.begin(), q.end());
    q.erase(std::unique(q.begin(), q.end()), q.end());
    TVZ_REQUIRE(static_cast<int>(q.size()) <= kFragMaxQ, "fragment query has %d distinct cuts; the limit is %d",
                static_cast<int>(q.size()), kFragMaxQ);
    fq.qn = static_cast<int>(q.size());
    fq.tol = tol;
    fq.tol_gap = tol_gap;
    fq.zero_only = zero_only;
    fq.exhaustive = anchor == 0 && !zero_only;
    std::copy(q.begin(), q.end(), fq.q);
    if (c->n_rows == 0) {
        TVZ_REQUIRE(!gather, "an empty shard cannot take part in the fused gather");
        TVZ_CUDA(cudaMemsetAsync(d_out, 0, 8, st));
        TVZ_CUDA(cudaMemsetAsync(c->d_nhits, 0, 8, st));
        return TVZ_OK;
    }
    int shift = 4;  // bucket width: a power of two above tol_gap, so a tolerance window spans <= 3 buckets
    while ((1ll << shift) <= tol_gap && shift < 30) ++shift;
    TVZ_REQUIRE((kGapRange >> shift) >= 1, "tol_gap %d too large", tol_gap);
    fq.min_match = min_match;
    if (!zero_only && anchor >= 2) {
        // streaming kernel; a query with fewer than `anchor` intervals generates no candidate at all
        if (fq.qn - 1 >= anchor) {
            const long long chunks = c->n_padded / kCtaChunk;
            // test hook: TVZ_FRAG_QUEUE_CAP=<n> (>= 0, < 64) forces early in-place drains
            int queue_cap = kStreamQueue;
            if (const char *env = getenv("TVZ_FRAG_QUEUE_CAP")) queue_cap = std::min(64, std::max(0, atoi(env)));
            static const int l2_ahead = [] {  // tuning hook: TVZ_FRAG_L2_AHEAD=<n> turns the L2 prefetch on
                const char *e = getenv("TVZ_FRAG_L2_AHEAD");
                return e ? atoi(e) : 0;  // measured: 0.115 ms without, 0.121 ms with (distance 1..4)
            }();
            const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(chunks, static_cast<long long>(kStreamMinBlocks) * num_sms())));
            auto kern = anchor == 2 ? fragment_stream_kernel<2> : fragment_stream_kernel<3>;
            TVZ_CUDA(launch_pdl(kern, dim3(grid), dim3(kStreamThreads), 0, st, c->d_ticks, c->n_vals, c->n_padded, c->d_off,
                                c->d_block_row, c->n_rows, fq, shift, c->d_keys, l2_ahead, queue_cap));
        }
        return compact_enqueue_keys(c->d_keys, c->n_rows, min_match, c->d_vid, d_out, c->d_rows, out_cap, c->d_nhits,
                                    c->d_state, c->d_ticket, d_out + 2 * (out_cap + 1), st, gather);
    }
    const long long want = (c->n_rows + kFragWarps - 1) / kFragWarps;
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(want, 4ll * num_sms())));
    fragment_match_kernel<<<grid, kFragThreads, 0, st>>>(c->d_ticks, c->d_off, c->n_rows, fq, shift, c->d_score,
                                                         c->d_delta);
    TVZ_CUDA(cudaGetLastError());
    // records at d_out[0 .. 2*(out_cap+1)), deltas behind them at d_out + 2*(out_cap+1)
    return compact_enqueue(c->d_score, c->n_rows, min_match, c->d_vid, d_out, c->d_rows, out_cap, c->d_nhits,
                           c->d_state, c->d_ticket, c->d_delta, d_out + 2 * (out_cap + 1), st, gather);
}

}  // namespace

extern "C" {

int tvz_fragcat_create(const double *h_ts, const int64_t *h_off, const int32_t *h_video_id, int64_t n_rows,
                       double tick_hz, int64_t hit_capacity, tvz_fragcat **out) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(out, "null out pointer");
    *out = nullptr;
    TVZ_REQUIRE(n_rows >= 0 && tick_hz > 0 && hit_capacity >= 0, "bad arguments");
    TVZ_REQUIRE(n_rows == 0 || (h_off && h_video_id && h_off[0] == 0), "bad offsets/video ids");
    std::vector<int> ticks;
    std::vector<long long> off(static_cast<size_t>(n_rows) + 1, 0);
    ticks.reserve(n_rows ? static_cast<size_t>(h_off[n_rows]) : 0);
    for (int64_t r = 0; r < n_rows; ++r) {
        TVZ_REQUIRE(h_off[r + 1] >= h_off[r], "offsets must be non-decreasing (row %lld)", (long long)r);
        const size_t start = ticks.size();
        bool sorted = true;
        for (int64_t j = h_off[r]; j < h_off[r + 1]; ++j) {
            int t;
            if (!to_tick(h_ts[j], tick_hz, &t)) continue;
            if (ticks.size() > start && t <= ticks.back()) sorted = false;
            ticks.push_back(t);
        }
        if (!sorted) {
            std::sort(ticks.begin() + start, ticks.end());
            ticks.erase(std::unique(ticks.begin() + start, ticks.end()), ticks.end());
        }
        off[r + 1] = static_cast<long long>(ticks.size());
    }
    tvz_fragcat *c = new tvz_fragcat();
    c->tick_hz = tick_hz;
    c->n_rows = n_rows;
    c->n_vals = static_cast<long long>(ticks.size());
    c->n_padded = (c->n_vals + kCtaChunk - 1) / kCtaChunk * kCtaChunk;
    ticks.resize(static_cast<size_t>(c->n_padded) + 8, kPadTick);
    // coarse index: last row whose offset is <= b * 256 (clamped to the last row)
    std::vector<int> block_row(static_cast<size_t>(c->n_padded >> kStreamBlockShift) + 2, 0);
    {
        long long r = 0;
        for (size_t b = 0; b < block_row.size(); ++b) {
            const long long e = static_cast<long long>(b) << kStreamBlockShift;
            while (r + 1 < n_rows && off[r + 1] <= e) ++r;
            block_row[b] = static_cast<int>(r);
        }
    }
    cudaGetDevice(&c->device);
    auto fail = [&](cudaError_t e, const char *what) {
        set_error(TVZ_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        tvz_fragcat_destroy(c);
        return TVZ_ERR_CUDA;
    };
    cudaError_t e;
    const size_t nr = std::max<long long>(1, n_rows);
    if ((e = cudaMalloc(&c->d_ticks, std::max<size_t>(1, ticks.size()) * 4)) != cudaSuccess) return fail(e, "cudaMalloc(ticks)");
    if ((e = cudaMalloc(&c->d_off, off.size() * 8)) != cudaSuccess) return fail(e, "cudaMalloc(off)");
    if ((e = cudaMalloc(&c->d_vid, nr * 4)) != cudaSuccess) return fail(e, "cudaMalloc(vid)");
    if ((e = cudaMalloc(&c->d_block_row, block_row.size() * 4)) != cudaSuccess) return fail(e, "cudaMalloc(block_row)");
    if ((e = cudaMemcpy(c->d_block_row, block_row.data(), block_row.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(block_row)");
    if ((e = cudaMalloc(&c->d_keys, nr * 8)) != cudaSuccess) return fail(e, "cudaMalloc(keys)");
    if ((e = cudaMemset(c->d_keys, 0, nr * 8)) != cudaSuccess) return fail(e, "cudaMemset(keys)");
    if ((e = cudaMalloc(&c->d_score, nr * 4)) != cudaSuccess) return fail(e, "cudaMalloc(score)");
    if ((e = cudaMalloc(&c->d_delta, nr * 4)) != cudaSuccess) return fail(e, "cudaMalloc(delta)");
    if ((e = cudaMalloc(&c->d_nhits, 8)) != cudaSuccess) return fail(e, "cudaMalloc(nhits)");
    const int nb = std::max(1, compact_blocks(n_rows));
    if ((e = cudaMalloc(&c->d_state, nb * 8)) != cudaSuccess) return fail(e, "cudaMalloc(state)");
    if ((e = cudaMalloc(&c->d_ticket, 12)) != cudaSuccess) return fail(e, "cudaMalloc(ticket)");
    if ((e = cudaMemset(c->d_state, 0, nb * 8)) != cudaSuccess) return fail(e, "cudaMemset(state)");
    if ((e = cudaMemset(c->d_score, 0, nr * 4)) != cudaSuccess) return fail(e, "cudaMemset(score)");
    if ((e = cudaMemset(c->d_nhits, 0, 8)) != cudaSuccess) return fail(e, "cudaMemset(nhits)");
    const unsigned init[3] = {0u, 1u, 0u};
    if ((e = cudaMemcpy(c->d_ticket, init, 12, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "cudaMemcpy(ticket)");
    if (!ticks.empty() &&
        (e = cudaMemcpy(c->d_ticks, ticks.data(), ticks.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(ticks)");
    if ((e = cudaMemcpy(c->d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(off)");
    if (n_rows && (e = cudaMemcpy(c->d_vid, h_video_id, n_rows * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e, "cudaMemcpy(vid)");
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if (frag_reserve(c, std::max<long long>(1, hit_capacity)) != TVZ_OK) {
        tvz_fragcat_destroy(c);
        return TVZ_ERR_CUDA;
    }
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail(e, "cudaDeviceSynchronize");
    *out = c;
    return TVZ_OK;
    });
}

void tvz_fragcat_destroy(tvz_fragcat *c) {
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    void *dev[] = {c->d_ticks, c->d_off, c->d_vid, c->d_score, c->d_delta, c->d_out, c->d_rows, c->d_nhits, c->d_state,
                   c->d_ticket, c->d_block_row, c->d_keys};
    for (void *p : dev)
        if (p) cudaFree(p);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int64_t tvz_fragcat_rows(const tvz_fragcat *c) { return c ? c->n_rows : 0; }
int64_t tvz_fragcat_values(const tvz_fragcat *c) { return c ? c->n_vals : 0; }

int tvz_fragcat_match_async(tvz_fragcat *c, const double *h_q, int qn, int min_match, int tol_ticks,
                            int tol_gap_ticks, int anchor_intervals, int zero_offset_only, int32_t *d_out,
                            int64_t out_cap, void *stream) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(c && d_out && out_cap >= 1, "bad arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = frag_reserve(c, out_cap);  // d_rows must hold out_cap row indices
    if (rc) return rc;
    return frag_enqueue(c, h_q, qn, min_match, tol_ticks, tol_gap_ticks, anchor_intervals, zero_offset_only, d_out,
                        out_cap, static_cast<cudaStream_t>(stream));
    });
}

int tvz_fragcat_match_gather_async(tvz_fragcat *c, const double *h_q, int qn, int min_match, int tol_ticks,
                                   int tol_gap_ticks, int anchor_intervals, int zero_offset_only, int n_peers,
                                   const uint64_t *peer_record, const uint64_t *peer_flag, const uint32_t *d_my_flags,
                                   int64_t out_cap, uint32_t epoch, void *stream) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(c && out_cap >= 1, "bad arguments");
    TVZ_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers, "n_peers %d outside [1, %d]", n_peers, kMaxPeers);
    TVZ_REQUIRE(peer_record && peer_flag && d_my_flags, "null pointer");
    TVZ_REQUIRE(c->n_rows > 0, "the fused gather needs a non-empty shard");
    GatherTargets gt;
    gt.n_peers = n_peers;
    gt.epoch = epoch;
    for (int p = 0; p < n_peers; ++p) {
        gt.record[p] = reinterpret_cast<int *>(static_cast<uintptr_t>(peer_record[p]));
        gt.flag[p] = reinterpret_cast<unsigned *>(static_cast<uintptr_t>(peer_flag[p]));
    }
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = frag_reserve(c, out_cap);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = frag_enqueue(c, h_q, qn, min_match, tol_ticks, tol_gap_ticks, anchor_intervals, zero_offset_only, c->d_out,
                      out_cap, st, &gt);
    if (rc) return rc;
    return gather_wait_enqueue(d_my_flags, n_peers, epoch, st);
    });
}

int tvz_fragcat_match(tvz_fragcat *c, const double *q, int qn, int min_match, int tol_ticks, int tol_gap_ticks,
                      int anchor_intervals, int zero_offset_only, int32_t *out_video_id, int32_t *out_score, int32_t *out_delta_ticks,
                      int64_t cap, int64_t *n_out) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(c && n_out, "null pointer");
    *n_out = 0;
    TVZ_REQUIRE(cap >= 0 && (cap == 0 || (out_video_id && out_score && out_delta_ticks)), "bad output buffers");
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = frag_enqueue(c, q, qn, min_match, tol_ticks, tol_gap_ticks, anchor_intervals, zero_offset_only, c->d_out,
                          c->cap, c->stream);
    if (rc) return rc;
    TVZ_CUDA(cudaMemcpyAsync(c->h_out, c->d_out, 8, cudaMemcpyDeviceToHost, c->stream));
    TVZ_CUDA(cudaStreamSynchronize(c->stream));
    long long n_hits = c->h_out[0];
    if (n_hits == 0x7fffffff) TVZ_CUDA(cudaMemcpy(&n_hits, c->d_nhits, 8, cudaMemcpyDeviceToHost));
    *n_out = n_hits;
    if (n_hits > c->cap) {  // grow and let the caller retry: the scores were consumed by the compaction
        rc = frag_reserve(c, n_hits);
        if (rc) return rc;
        return set_error(TVZ_ERR_OVERFLOW, "%lld rows qualify; capacity grown, run the query again", n_hits);
    }
    if (n_hits > cap)
        return set_error(TVZ_ERR_OVERFLOW, "%lld rows qualify; caller capacity %lld", n_hits, (long long)cap);
    if (n_hits > 0) {
        TVZ_CUDA(cudaMemcpyAsync(c->h_out + 2, c->d_out + 2, n_hits * 8, cudaMemcpyDeviceToHost, c->stream));
        TVZ_CUDA(cudaMemcpyAsync(c->h_out + 2 * (c->cap + 1) + 1, c->d_out + 2 * (c->cap + 1) + 1, n_hits * 4,
                                 cudaMemcpyDeviceToHost, c->stream));
        TVZ_CUDA(cudaStreamSynchronize(c->stream));
    }
    for (long long h = 0; h < n_hits; ++h) {
        out_video_id[h] = c->h_out[2 + 2 * h];
        out_score[h] = c->h_out[3 + 2 * h];
        out_delta_ticks[h] = c->h_out[2 * (c->cap + 1) + 1 + h];
    }
    return TVZ_OK;
    });
}

}  // extern "C"
