/*
 * tvidz_b200.h -- C ABI of the B200-native TVIDZ analysis hot path.
 *
 * The reference (infraheads/tvidz) has no FFI: its seams for this path are a
 * Python function (inspector/db.py:76 find_duplicates) and a subprocess text
 * protocol (inspector/app.py:202-232, the ffmpeg `select=gt(scene\,0.3),showinfo`
 * pipeline).  These entry points are what a ctypes binding inside the inspector
 * would call instead; INTEGRATION.md shows that binding.
 *
 * Conventions: every function returns 0 on success and a negative tvz_status on
 * failure; tvz_last_error() returns a thread-local message.  Nothing throws,
 * nothing frees caller memory.  Pointers prefixed d_ are device pointers on the
 * current CUDA device, h_ are host pointers.  `stream` is a cudaStream_t passed
 * as void* (NULL = default stream).  All calls are re-entrant; a tvz_catalog may
 * be matched from several host threads at once provided each thread passes its
 * own tvz_match_ws (the reference runs one analysis thread per upload,
 * app.py:43,472).
 */
#ifndef TVIDZ_B200_H
#define TVIDZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tvz_status {
    TVZ_OK = 0,
    TVZ_ERR_INVALID = -1,   /* bad argument */
    TVZ_ERR_CUDA = -2,      /* CUDA runtime error; see tvz_last_error() */
    TVZ_ERR_NOMEM = -3,
    TVZ_ERR_OVERFLOW = -4   /* result did not fit the caller's capacity; *n_out holds the size needed */
} tvz_status;

const char *tvz_last_error(void);
int tvz_abi_version(void);

/* ------------------------------------------------------------------ stage 1
 * Replaces the arithmetic of FFmpeg's `select` scene detector that the
 * reference launches at inspector/app.py:202-209.
 */

/* Luma byte-SAD between consecutive frames (FFmpeg scene_sad.c ff_scene_sad_c,
 * invoked per frame by f_select.c get_scene_score; call site app.py:206).
 *   d_luma : 8-bit luma, frame t of stream s at
 *            d_luma + s*stream_stride_bytes + t*frame_stride_bytes, rows
 *            pitch_bytes apart, `width` visible bytes per row.
 *   d_sad  : uint64 [n_streams][n_frames]; d_sad[s][0] = 0 and d_sad[s][t] is
 *            the SAD of frames t-1 and t.  Exact integers.
 * Layouts whose base, pitch and strides are 16-byte multiples (with pitch ==
 * width, or width % 16 == 0) take the TMA bulk-copy kernel; anything else takes
 * a slower generic CUDA kernel.  There is no CPU path.
 */
int tvz_sad_luma_u8(const uint8_t *d_luma, int n_streams, int n_frames, int width, int height,
                    int64_t pitch_bytes, int64_t frame_stride_bytes, int64_t stream_stride_bytes,
                    uint64_t *d_sad, void *stream);

/* The same for planes of 16-bit samples (FFmpeg scene_sad.c ff_scene_sad16_c, the path the select
 * filter takes for yuv420p10): `width` in samples, pitch and strides in bytes.  Follow with
 * tvz_scene_select(..., bitdepth = 10), which divides mafd by 2^(bitdepth-8) as f_select.c does. */
int tvz_sad_luma_u16(const uint16_t *d_luma, int n_streams, int n_frames, int width, int height,
                     int64_t pitch_bytes, int64_t frame_stride_bytes, int64_t stream_stride_bytes,
                     uint64_t *d_sad, void *stream);

/* Which kernel tvz_sad_luma_u8 would pick for this layout: 1 = TMA bulk, 0 = generic. */
int tvz_sad_luma_u8_path(const uint8_t *d_luma, int width, int height, int64_t pitch_bytes,
                         int64_t frame_stride_bytes, int64_t stream_stride_bytes);

/* Scene score and selection (f_select.c get_scene_score + select_frame with
 * expression gt(scene,threshold); libavutil av_clipf float rounding included):
 *   mafd = (double)sad/(w*h)/2^(bitdepth-8); diff = |mafd - prev_mafd|;
 *   score = (double)(float)clip(min(mafd,diff)/100, 0, 1); frame 0 scores 0.
 *   selected = score > threshold.
 * d_score / d_selected may be NULL.
 */
int tvz_scene_select(const uint64_t *d_sad, int n_streams, int n_frames, int width, int height,
                     int bitdepth, double threshold, double *d_score, uint8_t *d_selected,
                     void *stream);

/* End-to-end entry with HOST buffers: the call a binding makes when decoded
 * frames sit in (ideally pinned) host memory.  Frames are streamed to the device
 * in chunks over two copy/compute streams, scored, and the three result arrays
 * [n_streams][n_frames] are written back to host memory.  Any of h_sad, h_score,
 * h_selected may be NULL.  chunk_frames <= 0 picks a default.  bitdepth 8: h_luma holds bytes;
 * bitdepth 9..16: h_luma holds 16-bit samples (width in samples, pitch/strides in bytes).
 */
int tvz_scene_score_host(const uint8_t *h_luma, int n_streams, int n_frames, int width, int height,
                         int64_t pitch_bytes, int64_t frame_stride_bytes, int64_t stream_stride_bytes,
                         int bitdepth, double threshold, int chunk_frames,
                         uint64_t *h_sad, double *h_score, uint8_t *h_selected);

/* ------------------------------------------------------------------ decode front-end
 * Replaces the software decode inside the ffmpeg process of inspector/app.py:202-208: demuxed packets
 * go to the GPU's hardware decoder (NVDEC, through the driver's libnvcuvid.so.1, loaded at run time),
 * the luma plane of every decoded frame lands in a dense ring in device memory -- the plane FFmpeg's
 * scene filter reads for 4:2:0 sources -- and tvz_sad_luma_u8 / _u16 score it there: raw frames never
 * cross PCIe.  There is no software fallback: without the library every call fails with a message.
 */
typedef enum tvz_codec {
    TVZ_CODEC_MPEG2 = 1, TVZ_CODEC_MPEG4 = 2, TVZ_CODEC_H264 = 3, TVZ_CODEC_HEVC = 4,
    TVZ_CODEC_VP8 = 5, TVZ_CODEC_VP9 = 6, TVZ_CODEC_AV1 = 7
} tvz_codec;
typedef struct tvz_decoder tvz_decoder;
int tvz_nvdec_available(void);                                   /* 1, or 0 with tvz_last_error() saying why */
const char *tvz_nvdec_library(void);                             /* the libnvcuvid that was loaded ("" = none) */
/* out4 = { supported, max coded width, max coded height, NVDEC engines } for 4:2:0 at `bitdepth` */
int tvz_nvdec_caps(int codec, int bitdepth, int32_t *out4);
int tvz_decoder_create(int codec, int64_t ring_frames, tvz_decoder **out);
void tvz_decoder_destroy(tvz_decoder *dec);
/* One demuxed packet (a frame of the elementary stream; H.264/HEVC in Annex-B form) -> parser -> NVDEC.
 * size 0 with end_of_stream != 0 flushes.  *frames_total = frames complete in the ring so far (display
 * order); the caller consumes [consumed, *frames_total) before the ring wraps over them (one packet
 * adds at most the decoder's surface count, <= 24 frames). */
int tvz_decoder_feed(tvz_decoder *dec, const uint8_t *packet, int64_t size, int64_t pts, int end_of_stream,
                     int64_t *frames_total);
/* out6 = { width, height, bitdepth, ring_frames, frames decoded, decode surfaces }; width == 0 until the
 * first sequence header has been parsed.  The ring is [ring_frames][height][width] samples (uint8, or
 * uint16 with the sample in the high bits when bitdepth > 8), dense; frame n sits in slot n % ring_frames. */
int tvz_decoder_info(const tvz_decoder *dec, int64_t *out6);
const uint8_t *tvz_decoder_ring(const tvz_decoder *dec);          /* device pointer */
int tvz_decoder_pts(const tvz_decoder *dec, int64_t first, int64_t n, int64_t *out);

/* ------------------------------------------------------------------ stage 2
 * Replaces inspector/db.py:76-94 find_duplicates over the rows of
 * `video_timestamps` (db.py:22-27: one float8[] per video).
 */
typedef struct tvz_catalog tvz_catalog;     /* device-resident packed catalogue (one shard) */
typedef struct tvz_match_ws tvz_match_ws;   /* per-thread query workspace */

/* Pack rows (CSR: h_ts[h_off[r] .. h_off[r+1]) is row r, h_video_id[r] its
 * videos.id) onto the current device.  Stored values are canonicalised so that
 * bitwise equality equals Python float `==` (db.py:88): -0.0 -> +0.0, NaNs
 * dropped, repeats inside a row dropped (they never change a match count).
 * Device memory: 26 bytes per stored value (a 16-bit fingerprint -- what a query
 * streams --, a 16-byte verification record {value, row}, and the value in row
 * order for the early-exit kernel) + 13 bytes per row. */
int tvz_catalog_create(const double *h_ts, const int64_t *h_off, const int32_t *h_video_id,
                       int64_t n_rows, tvz_catalog **out);
/* The same with a mutable TAIL of `tail_values` stored values (and up to 4096 rows) behind the
 * packed rows, for tvz_catalog_upsert. */
int tvz_catalog_create_mutable(const double *h_ts, const int64_t *h_off, const int32_t *h_video_id,
                               int64_t n_rows, int64_t tail_values, tvz_catalog **out);
void tvz_catalog_destroy(tvz_catalog *cat);
int64_t tvz_catalog_rows(const tvz_catalog *cat);       /* packed rows + rows in the tail (replaced ones included) */
int64_t tvz_catalog_values(const tvz_catalog *cat);     /* stored doubles after canonicalisation */
int64_t tvz_catalog_algo_bytes(const tvz_catalog *cat); /* 8*values + 8*(rows+1), SURVEY.md 8d */
int tvz_catalog_tiles(const tvz_catalog *cat);          /* CTAs one query launches */

/* add_timestamps (db.py:43-64) on the device: the first live row of `video_id` is replaced
 * (its verification records are overwritten in place so that it can never match again), or a new
 * row is appended; the new row goes to the tail and is last in result order.  One small kernel
 * whose values ride in its parameters; queries enqueued after the call returns see it.
 * Contract: no query on this catalogue is in flight while an upsert runs (the Python Inspector
 * serialises them).  TVZ_ERR_OVERFLOW: the tail is full even after dropping replaced rows --
 * repack (create a new catalogue from the live rows). */
int tvz_catalog_upsert(tvz_catalog *cat, int32_t video_id, const double *h_ts, int n);
/* out4 = { rows in the tail, values in the tail, value capacity of the tail, replaced packed rows } */
int tvz_catalog_tail_info(const tvz_catalog *cat, int64_t *out4);

int tvz_match_ws_create(const tvz_catalog *cat, int64_t hit_capacity, tvz_match_ws **out);
void tvz_match_ws_destroy(tvz_match_ws *ws);

/* find_duplicates(new_timestamps=q[0..qn), min_match): every row whose
 * match_count = #{i : q[i] in row} is >= min_match, in catalogue order.
 * Host-buffer call: q and the outputs are host memory; the call returns after
 * the results are on the host.  out_kth (nullable) receives, per hit, the
 * 1-based query index at which the row reached min_match (0 if min_match <= 0):
 * the step at which the per-cut loop of app.py:231-255 first reports it. */
int tvz_catalog_match(const tvz_catalog *cat, tvz_match_ws *ws, const double *q, int qn, int min_match,
                      int32_t *out_video_id, int32_t *out_count, int32_t *out_kth,
                      int64_t cap, int64_t *n_out);

/* Batched find_duplicates: the reference runs one analysis thread per upload (app.py:43,472), so
 * concurrent queries are the normal case; up to tvz_catalog_batch_size() (8) of them are answered by
 * ONE pass over the fingerprints (one bit per query in the filter map, 16-bit counts), more are
 * processed group by group.
 *   q_all / q_off : CSR of the queries (q_off has n_queries + 1 entries)
 *   out_off       : int64 [n_queries + 1]; hits of query i are out_*[out_off[i] .. out_off[i+1])
 * Each query may hold at most tvz_catalog_batch_limit() distinct values and 65535 values in all.
 * On TVZ_ERR_OVERFLOW *need_per_query (if > the workspace's hit_capacity) and *need_total say how
 * much room a retry needs. */
int tvz_catalog_batch_limit(void);
int tvz_catalog_batch_size(void);
int tvz_catalog_match_batch(const tvz_catalog *cat, tvz_match_ws *ws, const double *q_all, const int64_t *q_off,
                            int n_queries, int min_match, int32_t *out_video_id, int32_t *out_count,
                            int64_t *out_off, int64_t cap_total, int64_t *need_per_query, int64_t *need_total);

/* Device-resident variant for pipelines and the sharded matcher: enqueues the
 * query -- ONE ordinary kernel launch that streams, counts and compacts (the keys
 * of a query with <= 224 distinct values travel in the kernel parameters) -- on
 * `stream` and returns without synchronising.  Back-to-back calls on one stream
 * pipeline without a host-side wait.  The result is written to
 *   d_out : int32 [out_cap + 1][2] on the device (NULL = the workspace's own
 *           buffer, see tvz_match_ws_hits, capacity = hit_capacity):
 *             d_out[0]     = { n_hits saturated to INT32_MAX, 1 if n_hits > out_cap }
 *             d_out[1 + h] = { video_id, match_count } of the h-th qualifying row
 *           -- one fixed-size record a collective can gather as is.
 *   tvz_match_ws_nhits(): int64, the exact number of qualifying rows.
 * out_cap must not exceed the workspace's hit_capacity.  h_q is consumed before
 * the call returns (it is staged into pinned memory owned by the workspace).
 */
int tvz_catalog_match_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *h_q, int qn,
                            int min_match, int32_t *d_out, int64_t out_cap, void *stream);
/* ... and up to 8 queries at once, answered by ONE kernel launch and one pass over the catalogue: d_out is
 * int32 [n_queries][out_cap + 1][2].  Each query: <= tvz_catalog_batch_limit() distinct values and <= 65535
 * values in all; the keys of all 8 travel in the kernel parameters (no copy in front of the launch), q_all is
 * consumed before the call returns. */
int tvz_catalog_match_batch_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *q_all,
                                  const int64_t *q_off, int n_queries, int min_match, int32_t *d_out,
                                  int64_t out_cap, void *stream);
/* Sharded matcher with the gather fused into the kernel (one process per GPU, peers reachable
 * over NVLink): like tvz_catalog_match_async into the workspace's own record, but every CTA also
 * STORES its hits (and the last tile the header) into this rank's slot of every peer's gather
 * buffer, in a TAGGED form: entry e of a slot is 16 bytes {value0, epoch, value1, epoch} -- entry 0 =
 * {n_hits, overflow}, entry 1 + h = {video_id, match_count} -- and every 8-byte half is stored
 * atomically, so a reader that finds `epoch` in both halves has the data.  Senders need no
 * system-scope fence, no flag and no counter; the kernel's last CTAs poll (bounded) the slots of all
 * peers in this rank's OWN buffer until they are complete for `epoch` -- no second kernel, no
 * collective.  When the kernel has completed, every shard's record is in d_my_slots.
 *   peer_record[p] : device address (peer memory) of THIS rank's slot inside peer p's buffer;
 *                    a slot holds 4 * (out_cap + 1) ints (x 8 for the batched form), 16-byte aligned.
 *                    n_dst = n_peers such addresses, or n_dst = 1: ONE multicast address of the slot
 *                    (NVSwitch replicates every store to all peers: 1/n_peers of the store instructions)
 *   d_my_slots     : this rank's own buffer: n_peers slots, slot_stride_ints apart (slot p is written by peer p)
 * Use a different buffer set for consecutive epochs (double buffering): a peer may start query k+1
 * while this rank still reads the records of query k.  Buffers start zeroed; epoch != 0.  n_peers <= 8. */
int tvz_catalog_match_gather_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *h_q, int qn,
                                   int min_match, int n_peers, int n_dst, const uint64_t *peer_record,
                                   const int32_t *d_my_slots, int64_t slot_stride_ints, int64_t out_cap,
                                   uint32_t epoch, void *stream);
/* The batched form: a slot is int32 [8][out_cap + 1][4]. */
int tvz_catalog_match_batch_gather_async(const tvz_catalog *cat, tvz_match_ws *ws, const double *q_all,
                                         const int64_t *q_off, int n_queries, int min_match, int n_peers, int n_dst,
                                         const uint64_t *peer_record, const int32_t *d_my_slots,
                                         int64_t slot_stride_ints, int64_t out_cap, uint32_t epoch,
                                         void *stream);
/* Strided device -> host copy of a slice of n fixed-size records: bytes [offset, offset + width) of
 * every record, pitch_bytes apart on the device and in h_rec alike; sync != 0 waits for the stream.
 * How the sharded matchers read every shard's header plus an optimistic first slice of every hit
 * list with a single device-to-host transfer. */
int tvz_copy_records_to_host(const void *d_rec, void *h_rec, int n_records, int64_t pitch_bytes,
                             int64_t offset_bytes, int64_t width_bytes, int sync, void *stream);
const int32_t *tvz_match_ws_hits(const tvz_match_ws *ws);
const int64_t *tvz_match_ws_nhits(const tvz_match_ws *ws);

/* ------------------------------------------------------------------ fragment mode
 * Offset-invariant matching of a clip's cut list against longer stored videos
 * (BASELINE.json config 5).  The reference advertises this (README.md:5) but
 * implements only offset-0 exact membership (inspector/db.py:78-79), so the
 * semantics are this library's own (tvidz_b200/csrc/fragment.cu header):
 *   ticks = llround(ts * tick_hz); candidate offsets d = C[j] - Q[i] come from
 *   positions where `anchor_intervals` (0..3) consecutive intervals of row and query agree
 *   within tol_gap ticks (0 = every pair (i, j): SURVEY.md B.4 as written, exhaustive and slow;
 *   1 = any agreeing adjacent pair; 2 = two in a row, far fewer candidates: the catalogue is
 *   then streamed once at HBM speed);
 *   score(d) = #{i : some C[j] within tol ticks of Q[i] + d}; a row reports its
 *   best (score, d) -- ties: smaller |d|, then smaller d -- iff score >= min_match.
 *   zero_offset_only = 1 scores d = 0 alone; with tol 0 that is find_duplicates'
 *   match_count (db.py:85-89) on tick-exact data.
 * Calls on one tvz_fragcat are serialised internally.
 */
typedef struct tvz_fragcat tvz_fragcat;
int tvz_fragcat_create(const double *h_ts, const int64_t *h_off, const int32_t *h_video_id, int64_t n_rows,
                       double tick_hz, int64_t hit_capacity, tvz_fragcat **out);
void tvz_fragcat_destroy(tvz_fragcat *cat);
int64_t tvz_fragcat_rows(const tvz_fragcat *cat);
int64_t tvz_fragcat_values(const tvz_fragcat *cat);   /* stored ticks (4 bytes each) */

/* Host-buffer call: results (catalogue order) in host arrays of `cap` entries. */
int tvz_fragcat_match(tvz_fragcat *cat, const double *q, int qn, int min_match, int tol_ticks, int tol_gap_ticks,
                      int anchor_intervals, int zero_offset_only, int32_t *out_video_id, int32_t *out_score, int32_t *out_delta_ticks,
                      int64_t cap, int64_t *n_out);

/* Device-resident variant: d_out is int32 [3 * (out_cap + 1)] on the device:
 *   d_out[0..1] = { n_hits saturated, overflow flag }, d_out[2 + 2h ..] = { video_id, score },
 *   d_out[2 * (out_cap + 1) + 1 + h] = offset ticks of hit h.  One fixed-size record per shard. */
int tvz_fragcat_match_async(tvz_fragcat *cat, const double *h_q, int qn, int min_match, int tol_ticks,
                            int tol_gap_ticks, int anchor_intervals, int zero_offset_only, int32_t *d_out,
                            int64_t out_cap, void *stream);

/* Sharded fragment matcher with the gather fused into the compaction kernel (see
 * tvz_catalog_match_gather_async): this rank's slot on every peer is int32 [3 * (out_cap + 1)] in the
 * layout above; a one-warp kernel behind the compaction waits for the peers' flags. */
int tvz_fragcat_match_gather_async(tvz_fragcat *cat, const double *h_q, int qn, int min_match, int tol_ticks,
                                   int tol_gap_ticks, int anchor_intervals, int zero_offset_only, int n_peers,
                                   const uint64_t *peer_record, const uint64_t *peer_flag,
                                   const uint32_t *d_my_flags, int64_t out_cap, uint32_t epoch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TVIDZ_B200_H */
