// tvz_scene_score_host: host buffers in, host results out.  What a binding inside the
// reference's analyze_file (inspector/app.py:197-232) calls with decoded luma planes in
// (pinned) host memory: frames are streamed host->device in chunks on a copy stream while
// the previous chunk is SAD-ed on a compute stream (double buffered; the last frame of a
// chunk is carried device-to-device into the next chunk so no byte crosses PCIe twice),
// then scored and copied back.
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace tvz {
int sad_accumulate(const uint8_t *d_luma, int n_streams, int n_frames, int width, int height, long long pitch,
                   long long frame_stride, long long stream_stride, unsigned long long *d_sad, long long sad_stride,
                   cudaStream_t st, bool is16);
int scene_select_launch(const unsigned long long *d_sad, int n_streams, int n_frames, int width, int height,
                        int bitdepth, double threshold, double *d_score, uint8_t *d_selected, cudaStream_t st);

namespace {

struct HostCtx {
    int device = -1;
    cudaStream_t copy = nullptr, compute = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr};    // the chunk in buffer b has landed
    cudaEvent_t released[2] = {nullptr, nullptr};  // buffer b may be overwritten
    uint8_t *buf[2] = {nullptr, nullptr};
    size_t buf_bytes = 0;
    unsigned long long *sad = nullptr;
    double *score = nullptr;
    uint8_t *sel = nullptr;
    size_t out_elems = 0;

    int init(int dev) {
        device = dev;
        TVZ_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        TVZ_CUDA(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            TVZ_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
            TVZ_CUDA(cudaEventCreateWithFlags(&released[i], cudaEventDisableTiming));
        }
        return TVZ_OK;
    }
    int reserve(size_t bytes_per_buf, size_t elems) {
        if (bytes_per_buf > buf_bytes) {
            for (int i = 0; i < 2; ++i) {
                if (buf[i]) cudaFree(buf[i]);
                buf[i] = nullptr;
            }
            buf_bytes = 0;
            for (int i = 0; i < 2; ++i) TVZ_CUDA(cudaMalloc(&buf[i], bytes_per_buf));
            buf_bytes = bytes_per_buf;
        }
        if (elems > out_elems) {
            if (sad) cudaFree(sad);
            if (score) cudaFree(score);
            if (sel) cudaFree(sel);
            sad = nullptr;
            score = nullptr;
            sel = nullptr;
            out_elems = 0;
            TVZ_CUDA(cudaMalloc(&sad, elems * sizeof(unsigned long long)));
            TVZ_CUDA(cudaMalloc(&score, elems * sizeof(double)));
            TVZ_CUDA(cudaMalloc(&sel, elems));
            out_elems = elems;
        }
        return TVZ_OK;
    }
};

std::mutex g_pool_mu;
std::vector<HostCtx *> g_pool;  // idle contexts; one is checked out per in-flight call (re-entrant)

HostCtx *checkout(int dev) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool.size(); ++i)
        if (g_pool[i]->device == dev) {
            HostCtx *c = g_pool[i];
            g_pool.erase(g_pool.begin() + i);
            return c;
        }
    return nullptr;
}
void checkin(HostCtx *c) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.push_back(c);
}

int run(HostCtx &cx, const uint8_t *h_luma, int S, int F, int W, int H, long long pitch, long long fstride,
        long long sstride, int bitdepth, double threshold, int C, uint64_t *h_sad, double *h_score,
        uint8_t *h_sel) {
    const size_t elems = static_cast<size_t>(S) * F;
    // Device chunk layout: [S][C+1] frames at the caller's frame stride.  Slot 0 holds the
    // carry (last frame of the previous chunk), slots 1..C the chunk's own frames.
    const long long d_sstride = static_cast<long long>(C + 1) * fstride;
    int rc = cx.reserve(static_cast<size_t>(S) * d_sstride, elems);
    if (rc) return rc;
    TVZ_CUDA(cudaMemsetAsync(cx.sad, 0, elems * sizeof(unsigned long long), cx.compute));
    const bool is16 = bitdepth > 8;
    const int Wb = is16 ? 2 * W : W;                                   // visible row bytes
    const size_t frame_span = static_cast<size_t>(H - 1) * pitch + Wb;  // bytes of one frame that matter
    int chunk = 0;
    for (int t0 = 0; t0 < F; t0 += C, ++chunk) {
        const int b = chunk & 1;
        const int n = std::min(F - t0, C);
        if (chunk >= 2) TVZ_CUDA(cudaStreamWaitEvent(cx.copy, cx.released[b], 0));
        const size_t bytes = static_cast<size_t>(n - 1) * fstride + frame_span;
        for (int s = 0; s < S; ++s)
            TVZ_CUDA(cudaMemcpyAsync(cx.buf[b] + s * d_sstride + fstride,
                                     h_luma + s * sstride + static_cast<long long>(t0) * fstride, bytes,
                                     cudaMemcpyHostToDevice, cx.copy));
        TVZ_CUDA(cudaEventRecord(cx.copied[b], cx.copy));
        if (chunk > 0) {
            // previous chunk was full: its last frame sits in slot C of the other buffer
            for (int s = 0; s < S; ++s)
                TVZ_CUDA(cudaMemcpyAsync(cx.buf[b] + s * d_sstride,
                                         cx.buf[b ^ 1] + s * d_sstride + static_cast<long long>(C) * fstride,
                                         frame_span, cudaMemcpyDeviceToDevice, cx.compute));
            TVZ_CUDA(cudaEventRecord(cx.released[b ^ 1], cx.compute));
        }
        TVZ_CUDA(cudaStreamWaitEvent(cx.compute, cx.copied[b], 0));
        if (chunk == 0)  // frames t0.. in slots 1..n, no predecessor
            rc = sad_accumulate(cx.buf[b] + fstride, S, n, Wb, H, pitch, fstride, d_sstride, cx.sad, F, cx.compute, is16);
        else  // slot 0 = frame t0-1: local index j <-> global frame t0-1+j
            rc = sad_accumulate(cx.buf[b], S, n + 1, Wb, H, pitch, fstride, d_sstride, cx.sad + (t0 - 1), F,
                                cx.compute, is16);
        if (rc) return rc;
    }
    rc = scene_select_launch(cx.sad, S, F, W, H, bitdepth, threshold, cx.score, cx.sel, cx.compute);
    if (rc) return rc;
    if (h_sad)
        TVZ_CUDA(cudaMemcpyAsync(h_sad, cx.sad, elems * sizeof(uint64_t), cudaMemcpyDeviceToHost, cx.compute));
    if (h_score)
        TVZ_CUDA(cudaMemcpyAsync(h_score, cx.score, elems * sizeof(double), cudaMemcpyDeviceToHost, cx.compute));
    if (h_sel) TVZ_CUDA(cudaMemcpyAsync(h_sel, cx.sel, elems, cudaMemcpyDeviceToHost, cx.compute));
    TVZ_CUDA(cudaStreamSynchronize(cx.compute));
    return TVZ_OK;
}

}  // namespace
}  // namespace tvz

using namespace tvz;

extern "C" int tvz_scene_score_host(const uint8_t *h_luma, int n_streams, int n_frames, int width, int height,
                                    int64_t pitch_bytes, int64_t frame_stride_bytes, int64_t stream_stride_bytes,
                                    int bitdepth, double threshold, int chunk_frames, uint64_t *h_sad,
                                    double *h_score, uint8_t *h_selected) {
    return guarded([&]() -> int {
    TVZ_REQUIRE(n_streams >= 0 && n_frames >= 0, "negative n_streams/n_frames");
    if (n_streams == 0 || n_frames == 0) return TVZ_OK;
    TVZ_REQUIRE(h_luma, "null frame pointer");
    TVZ_REQUIRE(width > 0 && height > 0, "width and height must be positive (got %dx%d)", width, height);
    TVZ_REQUIRE(bitdepth >= 8 && bitdepth <= 16, "bitdepth %d out of range (8..16)", bitdepth);
    const int64_t row_bytes = bitdepth > 8 ? 2 * (int64_t)width : width;
    TVZ_REQUIRE(pitch_bytes >= row_bytes, "pitch %lld < row of %lld bytes", (long long)pitch_bytes, (long long)row_bytes);
    TVZ_REQUIRE(frame_stride_bytes >= (int64_t)(height - 1) * pitch_bytes + row_bytes, "frames overlap");
    TVZ_REQUIRE(n_streams == 1 || stream_stride_bytes >= frame_stride_bytes * (int64_t)(n_frames - 1),
                "streams overlap");
    int C = chunk_frames;
    if (C <= 0) {
        const long long per_frame = static_cast<long long>(n_streams) * frame_stride_bytes;
        C = static_cast<int>(std::max<long long>(1, (512ll << 20) / std::max<long long>(1, per_frame)));
    }
    C = std::min(C, n_frames);
    int dev = 0;
    TVZ_CUDA(cudaGetDevice(&dev));
    HostCtx *cx = checkout(dev);
    if (!cx) {
        cx = new HostCtx();
        int rc = cx->init(dev);
        if (rc) return rc;  // leaked context on init failure: the process is unusable anyway
    }
    int rc = run(*cx, h_luma, n_streams, n_frames, width, height, pitch_bytes, frame_stride_bytes,
                 stream_stride_bytes, bitdepth, threshold, C, h_sad, h_score, h_selected);
    if (rc) cudaStreamSynchronize(cx->compute), cudaStreamSynchronize(cx->copy);
    checkin(cx);
    return rc;
    });
}
