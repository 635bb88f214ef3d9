// Stage 1: luma byte-SAD between consecutive frames + FFmpeg scene score.
//
// Replaces the arithmetic inside the ffmpeg `select=gt(scene\,0.3)` subprocess the
// reference launches at inspector/app.py:202-209 (FFmpeg libavfilter/scene_sad.c
// ff_scene_sad_c and f_select.c get_scene_score; restated in SURVEY.md App. A).
//
// Kernel design (HBM-bound, read-once):
//   * A work unit is (stream, time segment, spatial tile).  The CTA that owns a unit
//     walks the segment's frames in time order; the tile of frame t-1 stays in the
//     consumer threads' registers while the tile of frame t arrives, so every luma
//     byte crosses HBM exactly once (plus one carry frame per segment).
//   * One producer thread streams tiles global->shared with TMA 1-D bulk copies
//     (cp.async.bulk, SASS UBLKCP) through a kStages-deep ring guarded by full/empty
//     mbarriers; 8 consumer warps read the tile with LDS.128, SAD it against the
//     registers with VABSDIFF4.U8.ACC (__vsadu4), REDUX-reduce per warp and publish
//     one 64-bit RED per warp per (tile, frame) into sad[stream][frame].
//   * Tiles are either a contiguous byte range of a frame (pitch == width) or a band
//     of rows, one bulk copy per row (pitch > width, width % 16 == 0).
#include <algorithm>

#include "common.cuh"

namespace tvz {
namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kSadThreads = kConsumerThreads + 32;  // + one producer warp

struct SadParams {
    const uint8_t *base;
    unsigned long long *sad;   // [n_streams][sad_stride]
    long long sad_stride;      // elements between streams in `sad`
    int n_streams, n_frames;
    long long frame_stride, stream_stride;
    // tile geometry
    int n_tiles;          // tiles per frame
    int rows_mode;        // 0: contiguous byte ranges, 1: row bands
    int tile_units;       // flat: 16-byte units per (full) tile   | rows: rows per (full) tile
    int total_units;      // flat: 16-byte units per frame         | rows: rows per frame (height)
    int row_bytes;        // rows: visible bytes per row (width)
    long long pitch;      // rows: bytes between rows
    // time geometry
    int seg_len;          // SAD outputs per segment
    int n_segs;
    long long n_units;
};

struct Unit {
    int s, t0, nf;           // stream, first loaded frame, frames to load (nf-1 outputs)
    long long src_off;       // byte offset of the tile inside frame 0 of the stream
    int n_copies, copy_bytes;  // bulk copies per stage
    int n16;                 // 16-byte units in this tile
};

__device__ __forceinline__ Unit decode_unit(const SadParams &p, long long u) {
    Unit r;
    int k = static_cast<int>(u % p.n_tiles);
    long long sg = u / p.n_tiles;
    int g = static_cast<int>(sg % p.n_segs);
    r.s = static_cast<int>(sg / p.n_segs);
    r.t0 = g * p.seg_len;
    r.nf = min(p.seg_len, p.n_frames - 1 - r.t0) + 1;
    int first = k * p.tile_units;
    int cnt = min(p.tile_units, p.total_units - first);
    if (p.rows_mode) {
        r.src_off = static_cast<long long>(first) * p.pitch;
        r.n_copies = cnt;
        r.copy_bytes = p.row_bytes;
        r.n16 = cnt * (p.row_bytes >> 4);
    } else {
        r.src_off = static_cast<long long>(first) * 16;
        r.n_copies = 1;
        r.copy_bytes = cnt * 16;
        r.n16 = cnt;
    }
    return r;
}

template <int kVec>
__device__ __forceinline__ void load_tile(uint4 (&dst)[kVec], const uint4 *stage, int n16, int tid) {
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
        int idx = tid + j * kConsumerThreads;
        dst[j] = idx < n16 ? stage[idx] : make_uint4(0u, 0u, 0u, 0u);
    }
}

// k16 = false: 8-bit samples (ff_scene_sad_c);  k16 = true: 16-bit samples (ff_scene_sad16_c, the
// path FFmpeg takes for yuv420p10), two per 32-bit word.
template <int kVec, bool k16>
__device__ __forceinline__ unsigned sad_tile(const uint4 (&a)[kVec], const uint4 (&b)[kVec]) {
    unsigned acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;  // four chains: VABSDIFF4.ACC is a dependent add
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
        if (k16) {
            acc0 = __vsadu2(a[j].x, b[j].x) + acc0;
            acc1 = __vsadu2(a[j].y, b[j].y) + acc1;
            acc2 = __vsadu2(a[j].z, b[j].z) + acc2;
            acc3 = __vsadu2(a[j].w, b[j].w) + acc3;
        } else {
            acc0 = __vsadu4(a[j].x, b[j].x) + acc0;
            acc1 = __vsadu4(a[j].y, b[j].y) + acc1;
            acc2 = __vsadu4(a[j].z, b[j].z) + acc2;
            acc3 = __vsadu4(a[j].w, b[j].w) + acc3;
        }
    }
    return (acc0 + acc1) + (acc2 + acc3);  // <= 255 * 16 * kVec (u8) or 65535 * 8 * kVec (u16) per thread
}

template <int kStages, int kVec, bool k16>
__global__ void __launch_bounds__(kSadThreads) sad_bulk_kernel(const SadParams p) {
    constexpr int kTileCap = kConsumerThreads * kVec * 16;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + static_cast<size_t>(kStages) * kTileCap);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = smem_u32(bars + kStages);
    const uint32_t data0 = smem_u32(smem);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    int stage = 0;
    uint32_t phase = 0;

    if (warp == kConsumerWarps) {
        // ------------------------------------------------ TMA producer (one thread)
        if (lane == 0) {
            for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
                const Unit un = decode_unit(p, u);
                const uint8_t *src0 = p.base + un.s * p.stream_stride + un.src_off;
                const uint32_t bytes = static_cast<uint32_t>(un.n16) * 16u;
                for (int f = 0; f < un.nf; ++f) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                    const uint32_t full = full0 + 8 * stage;
                    mbar_arrive_expect_tx(full, bytes);
                    const uint8_t *src = src0 + static_cast<long long>(un.t0 + f) * p.frame_stride;
                    uint32_t dst = data0 + stage * kTileCap;
                    for (int c = 0; c < un.n_copies; ++c) {
                        tma_bulk_g2s(dst, src, static_cast<uint32_t>(un.copy_bytes), full);
                        dst += un.copy_bytes;
                        src += p.pitch;
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ---------------------------------------------------- consumers (8 warps)
    uint4 ra[kVec], rb[kVec];
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const Unit un = decode_unit(p, u);
        unsigned long long *out = p.sad + un.s * p.sad_stride + un.t0;

        auto acquire = [&](uint4(&dst)[kVec]) {
            mbar_wait(full0 + 8 * stage, phase);
            load_tile<kVec>(dst, reinterpret_cast<const uint4 *>(smem + static_cast<size_t>(stage) * kTileCap),
                            un.n16, tid);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * stage);  // tile is in registers: free the slot
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
        };
        auto publish = [&](unsigned v, int f) {
            v = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) atomicAdd(out + f, static_cast<unsigned long long>(v));
        };

        acquire(ra);
        int f = 1;
        for (; f + 1 < un.nf; f += 2) {
            acquire(rb);
            publish(sad_tile<kVec, k16>(ra, rb), f);
            acquire(ra);
            publish(sad_tile<kVec, k16>(rb, ra), f + 1);
        }
        if (f < un.nf) {
            acquire(rb);
            publish(sad_tile<kVec, k16>(ra, rb), f);
        }
    }
}

// Generic layouts (odd pitch, unaligned base, width % 16 != 0 with padding): plain
// loads, one CTA per (stream, frame pair, band of rows).  Correctness path only.
constexpr int kGenThreads = 256;
constexpr int kGenRows = 16;

__global__ void __launch_bounds__(kGenThreads) sad_generic_kernel(const uint8_t *__restrict__ base, int n_frames,
                                                                  int width, int height, long long pitch,
                                                                  long long frame_stride, long long stream_stride,
                                                                  unsigned long long *sad, long long sad_stride) {
    const int t = blockIdx.y + 1;
    const int s = blockIdx.z;
    if (t >= n_frames) return;
    const uint8_t *cur = base + s * stream_stride + t * frame_stride;
    const uint8_t *prv = cur - frame_stride;
    const int y0 = blockIdx.x * kGenRows;
    const int y1 = min(y0 + kGenRows, height);
    unsigned acc = 0;
    for (int y = y0; y < y1; ++y) {
        const uint8_t *a = prv + y * pitch;
        const uint8_t *b = cur + y * pitch;
        const bool same_align = ((reinterpret_cast<uintptr_t>(a) ^ reinterpret_cast<uintptr_t>(b)) & 3u) == 0;
        if (same_align) {
            int head = static_cast<int>((4u - (reinterpret_cast<uintptr_t>(a) & 3u)) & 3u);
            head = min(head, width);
            const int words = (width - head) >> 2;
            const int tail0 = head + words * 4;
            if (threadIdx.x < head) acc += abs(int(a[threadIdx.x]) - int(b[threadIdx.x]));
            const uint32_t *aw = reinterpret_cast<const uint32_t *>(a + head);
            const uint32_t *bw = reinterpret_cast<const uint32_t *>(b + head);
            for (int i = threadIdx.x; i < words; i += kGenThreads) acc = __vsadu4(aw[i], bw[i]) + acc;
            const int x = tail0 + threadIdx.x;
            if (x < width) acc += abs(int(a[x]) - int(b[x]));
        } else {
            for (int x = threadIdx.x; x < width; x += kGenThreads) acc += abs(int(a[x]) - int(b[x]));
        }
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    __shared__ unsigned warp_sums[kGenThreads / 32];
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        for (int i = 0; i < kGenThreads / 32; ++i) tot += warp_sums[i];
        atomicAdd(sad + s * sad_stride + t, tot);
    }
}

// 16-bit samples on layouts the bulk kernel cannot take: plain loads, same decomposition.
__global__ void __launch_bounds__(kGenThreads) sad_generic16_kernel(const uint8_t *__restrict__ base, int n_frames,
                                                                    int width, int height, long long pitch,
                                                                    long long frame_stride, long long stream_stride,
                                                                    unsigned long long *sad, long long sad_stride) {
    const int t = blockIdx.y + 1;
    const int s = blockIdx.z;
    if (t >= n_frames) return;
    const uint8_t *cur = base + s * stream_stride + t * frame_stride;
    const uint8_t *prv = cur - frame_stride;
    const int y0 = blockIdx.x * kGenRows;
    const int y1 = min(y0 + kGenRows, height);
    unsigned long long acc = 0;
    for (int y = y0; y < y1; ++y) {
        const uint16_t *a = reinterpret_cast<const uint16_t *>(prv + y * pitch);
        const uint16_t *b = reinterpret_cast<const uint16_t *>(cur + y * pitch);
        for (int x = threadIdx.x; x < width; x += kGenThreads) acc += abs(int(a[x]) - int(b[x]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ unsigned long long warp_sums[kGenThreads / 32];
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        for (int i = 0; i < kGenThreads / 32; ++i) tot += warp_sums[i];
        atomicAdd(sad + s * sad_stride + t, tot);
    }
}

// FFmpeg f_select.c get_scene_score + gt(scene, T): lag-1 dependency only, so every
// (stream, frame) is independent.  IEEE double division, one float32 rounding.
__global__ void scene_select_kernel(const unsigned long long *__restrict__ sad, int n_streams, int n_frames,
                                    double count, double depth_div, double threshold,
                                    double *__restrict__ score, uint8_t *__restrict__ selected) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long n = static_cast<long long>(n_streams) * n_frames;
    if (i >= n) return;
    const int t = static_cast<int>(i % n_frames);
    double ret = 0.0;
    if (t > 0) {
        const double mafd = __ddiv_rn(__ddiv_rn(static_cast<double>(sad[i]), count), depth_div);
        // prev_mafd: frame 0 never updates it (stays 0.0)
        const double prev = t > 1 ? __ddiv_rn(__ddiv_rn(static_cast<double>(sad[i - 1]), count), depth_div) : 0.0;
        const double diff = fabs(__dsub_rn(mafd, prev));
        const double m = mafd > diff ? diff : mafd;          // FFMIN(mafd, diff)
        float f = __double2float_rn(__ddiv_rn(m, 100.0));    // av_clipf takes a float
        f = f < 0.f ? 0.f : (f > 1.f ? 1.f : f);
        ret = static_cast<double>(f);
    }
    if (score) score[i] = ret;
    if (selected) selected[i] = ret > threshold ? 1 : 0;
}

// ---- tuning knobs (debug/sweep only; defaults chosen from profiles/) ----------------
struct SadTuning {
    int variant = 0;       // index into the variant table below
    int ctas_per_sm = 0;   // 0 = as many as shared memory allows (per variant)
    int units_per_cta = 16;  // target work units per CTA when splitting time into segments
    int min_seg = 16;
};
SadTuning g_tuning;

struct Variant {
    int stages, vec;
};
constexpr Variant kVariants[] = {{6, 8}, {4, 4}, {3, 8}, {8, 4}, {12, 4}, {4, 8}};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

template <int kStages, int kVec>
int launch_bulk(const SadParams &p, int ctas_per_sm, bool is16, cudaStream_t st) {
    constexpr int kTileCap = kConsumerThreads * kVec * 16;
    constexpr int smem = kStages * kTileCap + 2 * kStages * 8;
    static_assert(smem <= 227 * 1024, "ring does not fit shared memory");
    auto kern = is16 ? sad_bulk_kernel<kStages, kVec, true> : sad_bulk_kernel<kStages, kVec, false>;
    TVZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int fit = std::max(1, (227 * 1024) / (smem + 1024));
    if (ctas_per_sm <= 0 || ctas_per_sm > fit) ctas_per_sm = fit;
    long long grid = std::min<long long>(p.n_units, static_cast<long long>(num_sms()) * ctas_per_sm);
    kern<<<static_cast<unsigned>(grid), kSadThreads, smem, st>>>(p);
    TVZ_CUDA(cudaGetLastError());
    return TVZ_OK;
}

bool bulk_eligible(const uint8_t *d_luma, int width, int height, long long pitch, long long frame_stride,
                   long long stream_stride) {
    if ((reinterpret_cast<uintptr_t>(d_luma) & 15u) || (frame_stride & 15) || (stream_stride & 15)) return false;
    const long long plane = static_cast<long long>(width) * height;
    if (pitch == width && (plane & 15) == 0) return true;  // contiguous plane
    if ((width & 15) == 0 && (pitch & 15) == 0) return true;  // row bands
    return false;
}

}  // namespace

// d_sad must already be zero at [s][1..n_frames) (callers memset); accumulates SADs there.
// `width` is the visible row length in BYTES (= samples for 8-bit, 2 x samples for 16-bit).
int sad_accumulate(const uint8_t *d_luma, int n_streams, int n_frames, int width, int height, long long pitch,
                   long long frame_stride, long long stream_stride, unsigned long long *d_sad, long long sad_stride,
                   cudaStream_t st, bool is16) {
    if (n_streams <= 0 || n_frames <= 1) return TVZ_OK;
    if (!bulk_eligible(d_luma, width, height, pitch, frame_stride, stream_stride)) {
        dim3 grid((height + kGenRows - 1) / kGenRows, n_frames - 1, n_streams);
        TVZ_REQUIRE(grid.y <= 65535 && grid.z <= 65535,
                    "generic SAD path: n_frames-1 and n_streams must be <= 65535 (got %d, %d)", n_frames - 1,
                    n_streams);
        if (is16) {
            TVZ_REQUIRE(((reinterpret_cast<uintptr_t>(d_luma) | pitch | frame_stride | stream_stride) & 1) == 0,
                        "16-bit luma needs 2-byte aligned base, pitch and strides");
            sad_generic16_kernel<<<grid, kGenThreads, 0, st>>>(d_luma, n_frames, width / 2, height, pitch, frame_stride,
                                                               stream_stride, d_sad, sad_stride);
        } else {
            sad_generic_kernel<<<grid, kGenThreads, 0, st>>>(d_luma, n_frames, width, height, pitch, frame_stride,
                                                             stream_stride, d_sad, sad_stride);
        }
        TVZ_CUDA(cudaGetLastError());
        return TVZ_OK;
    }
    const SadTuning tn = g_tuning;
    const Variant v = kVariants[tn.variant];
    const int cap_units = kConsumerThreads * v.vec;  // 16-byte units per tile
    SadParams p{};
    p.base = d_luma;
    p.sad = d_sad;
    p.sad_stride = sad_stride;
    p.n_streams = n_streams;
    p.n_frames = n_frames;
    p.frame_stride = frame_stride;
    p.stream_stride = stream_stride;
    p.pitch = pitch;
    const long long plane = static_cast<long long>(width) * height;
    if (pitch == width && (plane & 15) == 0) {
        p.rows_mode = 0;
        const long long units = plane >> 4;
        TVZ_REQUIRE(units < (1ll << 31), "plane too large");
        p.total_units = static_cast<int>(units);
        p.n_tiles = static_cast<int>((units + cap_units - 1) / cap_units);
        p.tile_units = static_cast<int>((units + p.n_tiles - 1) / p.n_tiles);  // balanced tiles
        p.n_tiles = static_cast<int>((units + p.tile_units - 1) / p.tile_units);
    } else {
        p.rows_mode = 1;
        p.row_bytes = width;
        const int units_per_row = width >> 4;
        TVZ_REQUIRE(units_per_row <= cap_units, "row of %d bytes exceeds the %d-byte tile", width, cap_units * 16);
        int rows_cap = cap_units / units_per_row;
        p.total_units = height;
        p.n_tiles = (height + rows_cap - 1) / rows_cap;
        p.tile_units = (height + p.n_tiles - 1) / p.n_tiles;
        p.n_tiles = (height + p.tile_units - 1) / p.tile_units;
    }
    // Split time into segments only when there are too few (stream, tile) pairs to fill the GPU.
    const long long spatial = static_cast<long long>(n_streams) * p.n_tiles;
    const long long want = static_cast<long long>(num_sms()) * std::max(1, tn.units_per_cta);
    const int outputs = n_frames - 1;
    long long n_segs = (want + spatial - 1) / spatial;
    const long long max_segs = std::max(1, outputs / std::max(1, tn.min_seg));
    n_segs = std::max(1ll, std::min(n_segs, max_segs));
    p.seg_len = static_cast<int>((outputs + n_segs - 1) / n_segs);
    p.n_segs = (outputs + p.seg_len - 1) / p.seg_len;
    p.n_units = spatial * p.n_segs;

    switch (tn.variant) {
        case 0: return launch_bulk<6, 8>(p, tn.ctas_per_sm, is16, st);
        case 1: return launch_bulk<4, 4>(p, tn.ctas_per_sm, is16, st);
        case 2: return launch_bulk<3, 8>(p, tn.ctas_per_sm, is16, st);
        case 3: return launch_bulk<8, 4>(p, tn.ctas_per_sm, is16, st);
        case 4: return launch_bulk<12, 4>(p, tn.ctas_per_sm, is16, st);
        case 5: return launch_bulk<4, 8>(p, tn.ctas_per_sm, is16, st);
    }
    return set_error(TVZ_ERR_INVALID, "bad SAD variant %d", tn.variant);
}

int scene_select_launch(const unsigned long long *d_sad, int n_streams, int n_frames, int width, int height,
                        int bitdepth, double threshold, double *d_score, uint8_t *d_selected, cudaStream_t st) {
    const long long n = static_cast<long long>(n_streams) * n_frames;
    if (n <= 0) return TVZ_OK;
    const double count = static_cast<double>(static_cast<unsigned long long>(width) *
                                             static_cast<unsigned long long>(height));
    const double depth_div = static_cast<double>(1ull << (bitdepth - 8));
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    scene_select_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(d_sad, n_streams, n_frames, count,
                                                                          depth_div, threshold, d_score, d_selected);
    TVZ_CUDA(cudaGetLastError());
    return TVZ_OK;
}

}  // namespace tvz

using namespace tvz;

extern "C" {

int tvz_sad_luma_u8_path(const uint8_t *d_luma, int width, int height, int64_t pitch_bytes,
                         int64_t frame_stride_bytes, int64_t stream_stride_bytes) {
    return bulk_eligible(d_luma, width, height, pitch_bytes, frame_stride_bytes, stream_stride_bytes) ? 1 : 0;
}

int tvz_sad_luma_u8(const uint8_t *d_luma, int n_streams, int n_frames, int width, int height, int64_t pitch_bytes,
                    int64_t frame_stride_bytes, int64_t stream_stride_bytes, uint64_t *d_sad, void *stream) {
    TVZ_REQUIRE(n_streams >= 0 && n_frames >= 0, "negative n_streams/n_frames");
    if (n_streams == 0 || n_frames == 0) return TVZ_OK;
    TVZ_REQUIRE(d_luma && d_sad, "null pointer");
    TVZ_REQUIRE(width > 0 && height > 0, "width and height must be positive (got %dx%d)", width, height);
    TVZ_REQUIRE(pitch_bytes >= width, "pitch %lld < width %d", (long long)pitch_bytes, width);
    TVZ_REQUIRE(frame_stride_bytes >= 0 && stream_stride_bytes >= 0, "negative stride");
    TVZ_REQUIRE(static_cast<unsigned long long>(width) * height <= (1ull << 40), "frame too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TVZ_CUDA(cudaMemsetAsync(d_sad, 0, sizeof(uint64_t) * static_cast<size_t>(n_streams) * n_frames, st));
    return sad_accumulate(d_luma, n_streams, n_frames, width, height, pitch_bytes, frame_stride_bytes,
                          stream_stride_bytes, reinterpret_cast<unsigned long long *>(d_sad), n_frames, st, false);
}

int tvz_sad_luma_u16(const uint16_t *d_luma, int n_streams, int n_frames, int width, int height,
                     int64_t pitch_bytes, int64_t frame_stride_bytes, int64_t stream_stride_bytes, uint64_t *d_sad,
                     void *stream) {
    TVZ_REQUIRE(n_streams >= 0 && n_frames >= 0, "negative n_streams/n_frames");
    if (n_streams == 0 || n_frames == 0) return TVZ_OK;
    TVZ_REQUIRE(d_luma && d_sad, "null pointer");
    TVZ_REQUIRE(width > 0 && height > 0 && width < (1 << 29), "bad width/height (got %dx%d)", width, height);
    TVZ_REQUIRE(pitch_bytes >= 2 * (int64_t)width, "pitch %lld < 2 * width %d", (long long)pitch_bytes, width);
    TVZ_REQUIRE(frame_stride_bytes >= 0 && stream_stride_bytes >= 0, "negative stride");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TVZ_CUDA(cudaMemsetAsync(d_sad, 0, sizeof(uint64_t) * static_cast<size_t>(n_streams) * n_frames, st));
    return sad_accumulate(reinterpret_cast<const uint8_t *>(d_luma), n_streams, n_frames, 2 * width, height, pitch_bytes,
                          frame_stride_bytes, stream_stride_bytes, reinterpret_cast<unsigned long long *>(d_sad),
                          n_frames, st, true);
}

int tvz_scene_select(const uint64_t *d_sad, int n_streams, int n_frames, int width, int height, int bitdepth,
                     double threshold, double *d_score, uint8_t *d_selected, void *stream) {
    TVZ_REQUIRE(n_streams >= 0 && n_frames >= 0, "negative n_streams/n_frames");
    if (n_streams == 0 || n_frames == 0) return TVZ_OK;
    TVZ_REQUIRE(d_sad, "null pointer");
    TVZ_REQUIRE(width > 0 && height > 0, "width and height must be positive");
    TVZ_REQUIRE(bitdepth >= 8 && bitdepth <= 16, "bitdepth %d out of range", bitdepth);
    return scene_select_launch(reinterpret_cast<const unsigned long long *>(d_sad), n_streams, n_frames, width,
                               height, bitdepth, threshold, d_score, d_selected,
                               static_cast<cudaStream_t>(stream));
}

// Debug/sweep hook (not part of the public header): picks the ring geometry.
int tvz_debug_sad_tuning(int variant, int ctas_per_sm, int units_per_cta, int min_seg) {
    TVZ_REQUIRE(variant >= 0 && variant < kNumVariants, "variant out of range");
    g_tuning.variant = variant;
    g_tuning.ctas_per_sm = ctas_per_sm;
    if (units_per_cta > 0) g_tuning.units_per_cta = units_per_cta;
    if (min_seg > 0) g_tuning.min_seg = min_seg;
    return TVZ_OK;
}

}  // extern "C"
