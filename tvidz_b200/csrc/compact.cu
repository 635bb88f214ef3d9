// Single-pass ordered compaction of per-row results into a fixed-size hit record (fragment mode; the
// duplicate matcher compacts inside its own tile kernel, match.cu), plus the flag-wait kernel of the
// fused multi-GPU gather.
#include "common.cuh"

namespace tvz {
namespace {

// 16384 rows per block: the look-back walks its predecessors 32 at a time, and every hop is a dependent
// global round trip -- 1 M rows are 62 blocks (<= 2 hops) instead of 489 (<= 15 hops, ~8 us)
constexpr int kScanThreads = 1024;
constexpr int kScanRowsPerThread = 16;
constexpr int kScanRowsPerBlock = kScanThreads * kScanRowsPerThread;

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
                     unsigned long long *__restrict__ keys) {
    // ticket[0] = next ticket, ticket[1] = query epoch.  The epoch is read BEFORE the ticket is
    // taken and bumped by the holder of the last ticket, i.e. after every block has read it:
    // the kernel is self-contained and can be replayed from a CUDA graph.
    __shared__ unsigned s_block, s_epoch;
    __shared__ long long s_excl;
    __shared__ int ws[kScanThreads / 32];
    // Epoch and ticket are read AFTER the dependency wait: two compactions can be adjacent on a stream
    // (a query whose scan kernel is skipped), and the earlier one bumps the epoch / resets the ticket in
    // its last block.
    pdl_wait();  // everything before this kernel on the stream has finished
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        unsigned e;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(ticket + 1) : "memory");
        s_epoch = e;
        s_block = atomicAdd(ticket, 1u);
    }
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
            }
            ++pos;
        }
    }
    if (gt.n_peers == 0) return;

    // ---- fused gather epilogue: the block that finishes last ships the finished record to every peer ----
    // (one system-scope fence on the whole path, as in match_tile_kernel: header + hits, then the per-row
    // payload, which sits at the same distance behind the record on the peers as it does locally)
    __shared__ unsigned s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(ticket + 2, 1u);
        s_last = done == gridDim.x - 1;
        if (s_last) ticket[2] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        const int2 *src = reinterpret_cast<const int2 *>(out);
        const long long n_hits = min(static_cast<long long>(__ldcg(out)), cap);
        for (long long i = threadIdx.x; i < n_hits + 1; i += blockDim.x) {
            const int2 v = __ldcg(src + i);
            for (int p = 0; p < gt.n_peers; ++p) reinterpret_cast<int2 *>(gt.record[p])[i] = v;
        }
        if (aux_out) {
            const long long rel = aux_out - out;
            for (long long i = threadIdx.x; i < n_hits; i += blockDim.x) {
                const int v = __ldcg(aux_out + 1 + i);
                for (int p = 0; p < gt.n_peers; ++p) gt.record[p][rel + 1 + i] = v;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < gt.n_peers) {
        __threadfence_system();
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(gt.flag[threadIdx.x]), "r"(gt.epoch) : "memory");
    }
}

// Wait until every peer's record for `epoch` has landed in this rank's gather buffer.  Bounded:
// a peer that never answers turns into a launch failure, not a hung GPU.
__global__ void gather_wait_kernel(const unsigned *flags, int n_peers, unsigned epoch) {
    if (threadIdx.x >= n_peers) return;
    unsigned v, polls = 0;
    do {
        // (whatever reads the peers' records runs after this kernel: the poll needs no fence of its own)
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
        if (v == epoch) return;
        if (polls > 256) __nanosleep(64);
    } while (++polls < (1u << 26));
    __trap();
}

}  // namespace

int compact_blocks(long long n_rows) {
    return static_cast<int>((n_rows + kScanRowsPerBlock - 1) / kScanRowsPerBlock);
}

// Ordered compaction of counts[row] >= min_match (see match_compact_kernel).  `state` holds
// compact_blocks(n_rows) u64 records (zero-initialised once), `ticket` three u32 {0, 1, 0}.
int compact_enqueue(int *counts, long long n_rows, int min_match, const int *vid, int *out, long long *rows_out,
                    long long cap, long long *n_hits_out, unsigned long long *state, unsigned *ticket,
                    const int *aux, int *aux_out, cudaStream_t st, const GatherTargets *gather) {
    const GatherTargets none{};
    TVZ_CUDA(launch_pdl(match_compact_kernel<false>, dim3(compact_blocks(n_rows)), dim3(kScanThreads), 0, st, counts,
                        n_rows, min_match, vid, out, rows_out, cap, n_hits_out, state, ticket, aux, aux_out,
                        gather ? *gather : none, static_cast<unsigned long long *>(nullptr)));
    return TVZ_OK;
}

// Same compaction over packed best-candidate keys (fragment streaming kernel); keys[] is zeroed.
int compact_enqueue_keys(unsigned long long *keys, long long n_rows, int min_match, const int *vid, int *out,
                         long long *rows_out, long long cap, long long *n_hits_out, unsigned long long *state,
                         unsigned *ticket, int *delta_out, cudaStream_t st, const GatherTargets *gather) {
    const GatherTargets none{};
    TVZ_CUDA(launch_pdl(match_compact_kernel<true>, dim3(compact_blocks(n_rows)), dim3(kScanThreads), 0, st,
                        static_cast<int *>(nullptr), n_rows, min_match, vid, out, rows_out, cap, n_hits_out, state, ticket,
                        static_cast<const int *>(nullptr), delta_out, gather ? *gather : none, keys));
    return TVZ_OK;
}

int gather_wait_enqueue(const unsigned *d_flags, int n_peers, unsigned epoch, cudaStream_t st) {
    gather_wait_kernel<<<1, 32, 0, st>>>(d_flags, n_peers, epoch);
    TVZ_CUDA(cudaGetLastError());
    return TVZ_OK;
}

}  // namespace tvz
