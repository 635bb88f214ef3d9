// Decode front-end: compressed video -> luma planes in HBM through NVDEC.
//
// Replaces the software decode inside the ffmpeg process the reference launches (inspector/app.py:202-208):
// the elementary-stream packets of a file go to the GPU's hardware decoder, the luma plane of every
// decoded NV12 surface (same layout as yuv420p plane 0, FFmpeg's scene filter reads nothing else,
// SURVEY.md A.1) is copied device-to-device into a dense frame ring, and the SAD kernel reads it there.
// Raw frames never cross PCIe.
//
// libnvcuvid.so.1 ships with the driver (it is on the GPU boxes: gpurun probe) but neither the
// library nor its headers are in the build image, so it is dlopen()ed at run time and the handful of
// structures the parser-driven decode path needs are declared here (layout per NVIDIA Video Codec SDK
// nvcuvid.h / cuviddec.h; everything codec-specific stays opaque because the library's own parser fills it).
// If the library is absent every entry point fails with a clear message -- there is no software fallback.
#include <dlfcn.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace {

// ---- the slice of the nvcuvid ABI used here ----------------------------------------------------
typedef void *CUvideodecoder;
typedef void *CUvideoparser;
typedef long long CUvideotimestamp;
typedef int CUresultI;   // CUresult: 0 = success

enum { kCodecMPEG1 = 0, kCodecMPEG2, kCodecMPEG4, kCodecVC1, kCodecH264, kCodecJPEG, kCodecH264SVC, kCodecH264MVC,
       kCodecHEVC, kCodecVP8, kCodecVP9, kCodecAV1 };
enum { kSurfaceNV12 = 0, kSurfaceP016 = 1 };
enum { kChroma420 = 1 };
enum { kDeinterlaceWeave = 0 };
enum { kCreatePreferCUVID = 4 };
enum { kPktEndOfStream = 1, kPktTimestamp = 2, kPktEndOfPicture = 8 };

struct CUVIDDECODECAPS {
    int eCodecType, eChromaFormat;
    unsigned nBitDepthMinus8;
    unsigned reserved1[3];
    unsigned char bIsSupported, nNumNVDECs;
    unsigned short nOutputFormatMask;
    unsigned nMaxWidth, nMaxHeight, nMaxMBCount;
    unsigned short nMinWidth, nMinHeight;
    unsigned char bIsHistogramSupported, nCounterBitDepth;
    unsigned short nMaxHistogramBins;
    unsigned reserved3[10];
};

struct CUVIDDECODECREATEINFO {
    unsigned long ulWidth, ulHeight, ulNumDecodeSurfaces;
    int CodecType, ChromaFormat;
    unsigned long ulCreationFlags, bitDepthMinus8, ulIntraDecodeOnly, ulMaxWidth, ulMaxHeight, Reserved1;
    struct { short left, top, right, bottom; } display_area;
    int OutputFormat, DeinterlaceMode;
    unsigned long ulTargetWidth, ulTargetHeight, ulNumOutputSurfaces;
    void *vidLock;
    struct { short left, top, right, bottom; } target_rect;
    unsigned long enableHistogram;
    unsigned long Reserved2[4];
};
static_assert(sizeof(CUVIDDECODECREATEINFO) == 176, "CUVIDDECODECREATEINFO layout");

struct CUVIDEOFORMAT {
    int codec;
    struct { unsigned numerator, denominator; } frame_rate;
    unsigned char progressive_sequence, bit_depth_luma_minus8, bit_depth_chroma_minus8, min_num_decode_surfaces;
    unsigned coded_width, coded_height;
    struct { int left, top, right, bottom; } display_area;
    int chroma_format;
    unsigned bitrate;
    struct { int x, y; } display_aspect_ratio;
    struct { unsigned char bits, color_primaries, transfer_characteristics, matrix_coefficients; } video_signal_description;
    unsigned seqhdr_data_length;
};
static_assert(sizeof(CUVIDEOFORMAT) == 64, "CUVIDEOFORMAT layout");

struct CUVIDPICPARAMS_HEAD {   // only the head of CUVIDPICPARAMS is read; the pointer is passed through
    int PicWidthInMbs, FrameHeightInMbs, CurrPicIdx;
};

struct CUVIDPARSERDISPINFO {
    int picture_index, progressive_frame, top_field_first, repeat_first_field;
    CUvideotimestamp timestamp;
};

struct CUVIDSOURCEDATAPACKET {
    unsigned long flags, payload_size;
    const unsigned char *payload;
    CUvideotimestamp timestamp;
};

typedef int (*PFNSEQ)(void *, CUVIDEOFORMAT *);
typedef int (*PFNDEC)(void *, void *);
typedef int (*PFNDISP)(void *, CUVIDPARSERDISPINFO *);
struct CUVIDPARSERPARAMS {
    int CodecType;
    unsigned ulMaxNumDecodeSurfaces, ulClockRate, ulErrorThreshold, ulMaxDisplayDelay;
    unsigned bits;   // bAnnexb : 1, reserved : 31
    unsigned uReserved1[4];
    void *pUserData;
    PFNSEQ pfnSequenceCallback;
    PFNDEC pfnDecodePicture;
    PFNDISP pfnDisplayPicture;
    void *pfnGetOperatingPoint, *pfnGetSEIMsg;
    void *pvReserved2[5];
    void *pExtVideoInfo;
};
static_assert(sizeof(CUVIDPARSERPARAMS) == 136, "CUVIDPARSERPARAMS layout");

struct CUVIDPROCPARAMS {
    int progressive_frame, second_field, top_field_first, unpaired_field;
    unsigned reserved_flags, reserved_zero;
    unsigned long long raw_input_dptr;
    unsigned raw_input_pitch, raw_input_format;
    unsigned long long raw_output_dptr;
    unsigned raw_output_pitch, Reserved1;
    void *output_stream;
    unsigned Reserved[46];
    unsigned long long *histogram_dptr;
    void *Reserved2[1];
};
static_assert(sizeof(CUVIDPROCPARAMS) == 264, "CUVIDPROCPARAMS layout");

struct Api {
    void *handle = nullptr;
    CUresultI (*GetDecoderCaps)(CUVIDDECODECAPS *) = nullptr;
    CUresultI (*CreateDecoder)(CUvideodecoder *, CUVIDDECODECREATEINFO *) = nullptr;
    CUresultI (*DestroyDecoder)(CUvideodecoder) = nullptr;
    CUresultI (*DecodePicture)(CUvideodecoder, void *) = nullptr;
    CUresultI (*MapVideoFrame64)(CUvideodecoder, int, unsigned long long *, unsigned *, CUVIDPROCPARAMS *) = nullptr;
    CUresultI (*UnmapVideoFrame64)(CUvideodecoder, unsigned long long) = nullptr;
    CUresultI (*CreateVideoParser)(CUvideoparser *, CUVIDPARSERPARAMS *) = nullptr;
    CUresultI (*ParseVideoData)(CUvideoparser, CUVIDSOURCEDATAPACKET *) = nullptr;
    CUresultI (*DestroyVideoParser)(CUvideoparser) = nullptr;
    char why[256] = "";
    char path[256] = "";
};

Api *api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        // The copy that belongs to the RUNNING driver first (container runtimes mount it under
        // /usr/local/nvidia; an image may also carry a libnvcuvid of another driver version, which
        // answers CUDA_ERROR_NO_DEVICE to everything); TVZ_NVCUVID overrides.
        const char *env = getenv("TVZ_NVCUVID");
        const char *names[] = {env ? env : "", "/usr/local/nvidia/lib64/libnvcuvid.so.1", "/usr/local/nvidia/lib/libnvcuvid.so.1",
                               "libnvcuvid.so.1", "/usr/lib/x86_64-linux-gnu/libnvcuvid.so.1", "/usr/lib/libnvcuvid.so.1",
                               "libnvcuvid.so"};
        for (const char *n : names) {
            if (!n[0]) continue;
            a.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (a.handle) {
                snprintf(a.path, sizeof a.path, "%s", n);
                break;
            }
        }
        if (!a.handle) {
            snprintf(a.why, sizeof a.why, "libnvcuvid.so.1 not found (%s)", dlerror());
            return;
        }
        bool ok = true;
        auto sym = [&](const char *n) {
            void *p = dlsym(a.handle, n);
            if (!p) {
                ok = false;
                snprintf(a.why, sizeof a.why, "libnvcuvid has no symbol %s", n);
            }
            return p;
        };
        a.GetDecoderCaps = reinterpret_cast<decltype(a.GetDecoderCaps)>(sym("cuvidGetDecoderCaps"));
        a.CreateDecoder = reinterpret_cast<decltype(a.CreateDecoder)>(sym("cuvidCreateDecoder"));
        a.DestroyDecoder = reinterpret_cast<decltype(a.DestroyDecoder)>(sym("cuvidDestroyDecoder"));
        a.DecodePicture = reinterpret_cast<decltype(a.DecodePicture)>(sym("cuvidDecodePicture"));
        a.MapVideoFrame64 = reinterpret_cast<decltype(a.MapVideoFrame64)>(sym("cuvidMapVideoFrame64"));
        a.UnmapVideoFrame64 = reinterpret_cast<decltype(a.UnmapVideoFrame64)>(sym("cuvidUnmapVideoFrame64"));
        a.CreateVideoParser = reinterpret_cast<decltype(a.CreateVideoParser)>(sym("cuvidCreateVideoParser"));
        a.ParseVideoData = reinterpret_cast<decltype(a.ParseVideoData)>(sym("cuvidParseVideoData"));
        a.DestroyVideoParser = reinterpret_cast<decltype(a.DestroyVideoParser)>(sym("cuvidDestroyVideoParser"));
        if (!ok) {
            dlclose(a.handle);
            a.handle = nullptr;
        }
    });
    return &a;
}

int codec_id(int tvz_codec) {
    switch (tvz_codec) {
        case TVZ_CODEC_MPEG2: return kCodecMPEG2;
        case TVZ_CODEC_MPEG4: return kCodecMPEG4;
        case TVZ_CODEC_H264: return kCodecH264;
        case TVZ_CODEC_HEVC: return kCodecHEVC;
        case TVZ_CODEC_VP8: return kCodecVP8;
        case TVZ_CODEC_VP9: return kCodecVP9;
        case TVZ_CODEC_AV1: return kCodecAV1;
        default: return -1;
    }
}

}  // namespace

struct tvz_decoder {
    int codec = 0;
    CUvideoparser parser = nullptr;
    CUvideodecoder decoder = nullptr;
    cudaStream_t stream = nullptr;
    int width = 0, height = 0, bitdepth = 8;      // displayed size
    int surfaces = 0;
    long long ring_frames = 0;                    // frames the luma ring holds
    uint8_t *d_ring = nullptr;                    // [ring_frames][height][width] samples (1 or 2 bytes), dense
    bool own_ring = false;
    long long decoded = 0;                        // frames written to the ring so far (display order)
    std::vector<long long> pts;                   // their timestamps, as fed
    int error = 0;
    char why[256] = "";
    long long frame_bytes() const { return static_cast<long long>(width) * height * (bitdepth > 8 ? 2 : 1); }
};

namespace {

int fail(tvz_decoder *d, const char *what, int code) {
    if (!d->error) {
        d->error = TVZ_ERR_CUDA;
        snprintf(d->why, sizeof d->why, "%s failed (CUresult %d)", what, code);
    }
    return 0;   // stops the parser
}

int on_sequence(void *user, CUVIDEOFORMAT *f) {
    tvz_decoder *d = static_cast<tvz_decoder *>(user);
    Api *a = api();
    const int w = f->display_area.right - f->display_area.left, h = f->display_area.bottom - f->display_area.top;
    if (d->decoder) {
        if (w == d->width && h == d->height) return d->surfaces;   // repeated sequence header
        d->error = TVZ_ERR_INVALID;
        snprintf(d->why, sizeof d->why, "frame size changes mid-stream (%dx%d -> %dx%d): not supported", d->width, d->height, w, h);
        return 0;
    }
    if (f->chroma_format != kChroma420) {
        d->error = TVZ_ERR_INVALID;
        snprintf(d->why, sizeof d->why, "chroma format %d: only 4:2:0 sources are decoded here", f->chroma_format);
        return 0;
    }
    CUVIDDECODECAPS caps{};
    caps.eCodecType = f->codec;
    caps.eChromaFormat = f->chroma_format;
    caps.nBitDepthMinus8 = f->bit_depth_luma_minus8;
    int rc = a->GetDecoderCaps(&caps);
    if (rc) return fail(d, "cuvidGetDecoderCaps", rc);
    if (!caps.bIsSupported || f->coded_width > caps.nMaxWidth || f->coded_height > caps.nMaxHeight) {
        d->error = TVZ_ERR_INVALID;
        snprintf(d->why, sizeof d->why, "this GPU's NVDEC does not decode codec %d at %ux%u, %d bits", f->codec,
                 f->coded_width, f->coded_height, 8 + f->bit_depth_luma_minus8);
        return 0;
    }
    d->width = w;
    d->height = h;
    d->bitdepth = 8 + f->bit_depth_luma_minus8;
    d->surfaces = std::max<int>(f->min_num_decode_surfaces, 4) + 4;
    CUVIDDECODECREATEINFO ci{};
    ci.ulWidth = f->coded_width;
    ci.ulHeight = f->coded_height;
    ci.ulNumDecodeSurfaces = d->surfaces;
    ci.CodecType = f->codec;
    ci.ChromaFormat = f->chroma_format;
    ci.ulCreationFlags = kCreatePreferCUVID;
    ci.bitDepthMinus8 = f->bit_depth_luma_minus8;
    ci.ulMaxWidth = f->coded_width;
    ci.ulMaxHeight = f->coded_height;
    ci.display_area.left = static_cast<short>(f->display_area.left);
    ci.display_area.top = static_cast<short>(f->display_area.top);
    ci.display_area.right = static_cast<short>(f->display_area.right);
    ci.display_area.bottom = static_cast<short>(f->display_area.bottom);
    ci.OutputFormat = d->bitdepth > 8 ? kSurfaceP016 : kSurfaceNV12;
    ci.DeinterlaceMode = kDeinterlaceWeave;
    ci.ulTargetWidth = w;
    ci.ulTargetHeight = h;
    ci.ulNumOutputSurfaces = 2;
    rc = a->CreateDecoder(&d->decoder, &ci);
    if (rc) return fail(d, "cuvidCreateDecoder", rc);
    if (!d->d_ring) {
        if (cudaMalloc(&d->d_ring, d->ring_frames * d->frame_bytes()) != cudaSuccess) {
            cudaGetLastError();
            d->error = TVZ_ERR_NOMEM;
            snprintf(d->why, sizeof d->why, "cudaMalloc of the %lld-frame luma ring failed", d->ring_frames);
            return 0;
        }
        d->own_ring = true;
    }
    return d->surfaces;
}

int on_decode(void *user, void *pic) {
    tvz_decoder *d = static_cast<tvz_decoder *>(user);
    if (!d->decoder) return 0;
    const int rc = api()->DecodePicture(d->decoder, pic);
    if (rc) return fail(d, "cuvidDecodePicture", rc);
    return 1;
}

int on_display(void *user, CUVIDPARSERDISPINFO *info) {
    tvz_decoder *d = static_cast<tvz_decoder *>(user);
    if (!info || !d->decoder) return 1;   // (a null info marks the end of the stream)
    Api *a = api();
    CUVIDPROCPARAMS vpp{};
    vpp.progressive_frame = info->progressive_frame;
    vpp.second_field = info->repeat_first_field + 1;
    vpp.top_field_first = info->top_field_first;
    vpp.unpaired_field = info->repeat_first_field < 0;
    vpp.output_stream = d->stream;
    unsigned long long src = 0;
    unsigned pitch = 0;
    int rc = a->MapVideoFrame64(d->decoder, info->picture_index, &src, &pitch, &vpp);
    if (rc) return fail(d, "cuvidMapVideoFrame64", rc);
    // plane 0 of the NV12 / P016 surface: `height` rows, `pitch` bytes apart -> one dense frame of the ring
    const long long row = static_cast<long long>(d->width) * (d->bitdepth > 8 ? 2 : 1);
    uint8_t *dst = d->d_ring + (d->decoded % d->ring_frames) * d->frame_bytes();
    cudaError_t e = cudaMemcpy2DAsync(dst, row, reinterpret_cast<const void *>(src), pitch, row, d->height,
                                      cudaMemcpyDeviceToDevice, d->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);   // the surface goes back to the decoder below
    rc = a->UnmapVideoFrame64(d->decoder, src);
    if (e != cudaSuccess) {
        d->error = TVZ_ERR_CUDA;
        snprintf(d->why, sizeof d->why, "copy out of the decode surface failed: %s", cudaGetErrorString(e));
        return 0;
    }
    if (rc) return fail(d, "cuvidUnmapVideoFrame64", rc);
    d->pts.push_back(info->timestamp);
    ++d->decoded;
    return 1;
}

}  // namespace

extern "C" {

const char *tvz_nvdec_library(void) { return api()->path; }   /* which libnvcuvid was loaded ("" = none) */

int tvz_nvdec_available(void) {
    Api *a = api();
    if (a->handle) return 1;
    tvz::set_error(TVZ_ERR_INVALID, "%s", a->why);
    return 0;
}

/* out4 = { supported, max coded width, max coded height, NVDEC engines that serve it } for 4:2:0 at bitdepth */
int tvz_nvdec_caps(int codec, int bitdepth, int32_t *out4) {
    TVZ_REQUIRE(out4, "null pointer");
    Api *a = api();
    TVZ_REQUIRE(a->handle, "NVDEC is not available: %s", a->why);
    TVZ_REQUIRE(codec_id(codec) >= 0, "unknown codec %d", codec);
    TVZ_CUDA(cudaFree(nullptr));   // the runtime's context must be current: cuvid works on the current context
    CUVIDDECODECAPS caps{};
    caps.eCodecType = codec_id(codec);
    caps.eChromaFormat = kChroma420;
    caps.nBitDepthMinus8 = bitdepth - 8;
    const int rc = a->GetDecoderCaps(&caps);
    if (rc) return tvz::set_error(TVZ_ERR_CUDA, "cuvidGetDecoderCaps failed (CUresult %d)", rc);
    out4[0] = caps.bIsSupported;
    out4[1] = static_cast<int>(caps.nMaxWidth);
    out4[2] = static_cast<int>(caps.nMaxHeight);
    out4[3] = caps.nNumNVDECs;
    return TVZ_OK;
}

int tvz_decoder_create(int codec, int64_t ring_frames, tvz_decoder **out) {
    return tvz::guarded([&]() -> int {
    TVZ_REQUIRE(out, "null out pointer");
    *out = nullptr;
    Api *a = api();
    TVZ_REQUIRE(a->handle, "NVDEC is not available: %s (there is no software decode path in this library)", a->why);
    TVZ_REQUIRE(codec_id(codec) >= 0, "unknown codec %d", codec);
    TVZ_REQUIRE(ring_frames >= 8, "the luma ring needs at least 8 frames");
    TVZ_CUDA(cudaFree(nullptr));
    tvz_decoder *d = new tvz_decoder();
    d->codec = codec;
    d->ring_frames = ring_frames;
    if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete d;
        return tvz::set_error(TVZ_ERR_CUDA, "cudaStreamCreate failed");
    }
    CUVIDPARSERPARAMS pp{};
    pp.CodecType = codec_id(codec);
    pp.ulMaxNumDecodeSurfaces = 1;    // the sequence callback returns the real number
    pp.ulClockRate = 0;
    pp.ulMaxDisplayDelay = 2;         // lets decode run ahead of display
    pp.pUserData = d;
    pp.pfnSequenceCallback = on_sequence;
    pp.pfnDecodePicture = on_decode;
    pp.pfnDisplayPicture = on_display;
    const int rc = a->CreateVideoParser(&d->parser, &pp);
    if (rc) {
        cudaStreamDestroy(d->stream);
        delete d;
        return tvz::set_error(TVZ_ERR_CUDA, "cuvidCreateVideoParser failed (CUresult %d)", rc);
    }
    *out = d;
    return TVZ_OK;
    });
}

void tvz_decoder_destroy(tvz_decoder *d) {
    if (!d) return;
    Api *a = api();
    if (d->parser) a->DestroyVideoParser(d->parser);
    if (d->decoder) a->DestroyDecoder(d->decoder);
    if (d->own_ring && d->d_ring) cudaFree(d->d_ring);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

/* One demuxed packet (a frame or superframe of the elementary stream) -> parser -> NVDEC; size 0 with
 * end_of_stream != 0 flushes.  On return *frames_total = frames that have landed in the luma ring so far
 * (they are complete: the copy out of the decode surface is waited for).  The caller consumes frames
 * [consumed, *frames_total) from tvz_decoder_ring() before feeding so much that the ring wraps over them
 * (one packet adds at most the decoder's surface count, <= 24). */
int tvz_decoder_feed(tvz_decoder *d, const uint8_t *packet, int64_t size, int64_t pts, int end_of_stream,
                     int64_t *frames_total) {
    return tvz::guarded([&]() -> int {
    TVZ_REQUIRE(d && frames_total && size >= 0 && (size == 0 || packet), "bad arguments");
    if (d->error) return tvz::set_error(d->error, "%s", d->why);
    CUVIDSOURCEDATAPACKET pkt{};
    pkt.flags = kPktTimestamp | (size ? kPktEndOfPicture : 0) | (end_of_stream ? kPktEndOfStream : 0);
    pkt.payload_size = static_cast<unsigned long>(size);
    pkt.payload = packet;
    pkt.timestamp = pts;
    const int rc = api()->ParseVideoData(d->parser, &pkt);
    if (d->error) return tvz::set_error(d->error, "%s", d->why);
    if (rc) return tvz::set_error(TVZ_ERR_CUDA, "cuvidParseVideoData failed (CUresult %d)", rc);
    *frames_total = d->decoded;
    return TVZ_OK;
    });
}

/* out6 = { width, height, bitdepth, ring_frames, frames decoded, decode surfaces }; the ring is
 * [ring_frames][height][width] samples (uint8, or uint16 when bitdepth > 8), dense; frame n sits in slot
 * n % ring_frames.  width == 0 until the first sequence header has been parsed. */
int tvz_decoder_info(const tvz_decoder *d, int64_t *out6) {
    TVZ_REQUIRE(d && out6, "null pointer");
    out6[0] = d->width;
    out6[1] = d->height;
    out6[2] = d->bitdepth;
    out6[3] = d->ring_frames;
    out6[4] = d->decoded;
    out6[5] = d->surfaces;
    return TVZ_OK;
}
const uint8_t *tvz_decoder_ring(const tvz_decoder *d) { return d ? d->d_ring : nullptr; }
/* timestamps (as fed) of frames [first, first + n) in display order */
int tvz_decoder_pts(const tvz_decoder *d, int64_t first, int64_t n, int64_t *out) {
    TVZ_REQUIRE(d && out && first >= 0 && n >= 0 && first + n <= static_cast<int64_t>(d->pts.size()), "bad range");
    for (int64_t i = 0; i < n; ++i) out[i] = d->pts[first + i];
    return TVZ_OK;
}

}  // extern "C"
