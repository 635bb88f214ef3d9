/*
 * oracle/scene_oracle.c -- CPU restatement of FFmpeg's `select='gt(scene,T)'`
 * scene-change score, the arithmetic the reference reaches through the ffmpeg
 * subprocess it spawns at inspector/app.py:202-209.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under tvidz_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / CPU arm.
 *
 * PARITY UNPINNED: the arithmetic lives in FFmpeg (libavfilter/f_select.c
 * get_scene_score, libavfilter/scene_sad.c ff_scene_sad_c, libavutil/common.h
 * av_clipf, libavutil/timestamp.h av_ts_make_time_string), a binary dependency
 * the reference does not vendor and does not pin (inspector/Dockerfile:13 is a
 * bare `apt-get install ffmpeg`).  No ffmpeg binary or source exists in this
 * image and the reference's tests hold no scene-score vector,
 * so this file restates the published algorithm and is
 * pinned by the known-answer tests of Appendix A.5 in tests/ -- plus, for the
 * three pieces that live in libavutil, by the REAL library (FFmpeg 8.0.1's
 * libavutil 60.8, bundled with the image's OpenCV wheel): the pts_time text
 * against av_ts_make_time_string2, the gt(scene,T) verdict against
 * av_expr_parse_and_eval, byte-SADs against av_pixelutils block SADs
 * (tests/test_ffmpeg_libs.py).  get_scene_score itself (mafd, |d mafd|, the
 * float clip) has no reference-derived vector: for it parity stays unpinned.
 *
 * Each function names the upstream routine it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ff_scene_sad_c (libavfilter/scene_sad.c): byte SAD over the visible
 * width x height of one 8-bit plane; stride padding never contributes.
 * Call site in the reference: the `select` filter named at app.py:206. */
uint64_t tvzo_scene_sad_u8(const uint8_t *prev, int64_t prev_pitch,
                           const uint8_t *cur, int64_t cur_pitch,
                           int width, int height)
{
    uint64_t sad = 0;
    for (int y = 0; y < height; y++) {
        const uint8_t *a = prev + (int64_t)y * prev_pitch;
        const uint8_t *b = cur + (int64_t)y * cur_pitch;
        uint32_t row = 0; /* <= 255 * width, fits for width < 16.8M */
        for (int x = 0; x < width; x++) {
            int d = (int)a[x] - (int)b[x];
            row += (uint32_t)(d < 0 ? -d : d);
        }
        sad += row;
    }
    return sad;
}

/* ff_scene_sad16_c (libavfilter/scene_sad.c): the same over 16-bit samples -- the routine the
 * select filter installs when bitdepth > 8 (yuv420p10).  width in samples, pitches in bytes. */
uint64_t tvzo_scene_sad_u16(const uint8_t *prev, int64_t prev_pitch, const uint8_t *cur, int64_t cur_pitch,
                            int width, int height)
{
    uint64_t sad = 0;
    for (int y = 0; y < height; y++) {
        const uint16_t *a = (const uint16_t *)(prev + (int64_t)y * prev_pitch);
        const uint16_t *b = (const uint16_t *)(cur + (int64_t)y * cur_pitch);
        for (int x = 0; x < width; x++) {
            int d = (int)a[x] - (int)b[x];
            sad += (uint64_t)(d < 0 ? -d : d);
        }
    }
    return sad;
}

/* av_clipf_c (libavutil/common.h): float in, float out. */
static float clipf(float a, float amin, float amax)
{
    if (a < amin) return amin;
    else if (a > amax) return amax;
    else return a;
}

/* get_scene_score (libavfilter/f_select.c) applied to a whole stream whose
 * per-frame SADs are known: sad[0] is ignored (frame 0 has no predecessor:
 * score 0, prev_mafd stays 0).  select_frame then keeps frame t iff
 * gt(scene, threshold), i.e. (double)score > threshold. */
void tvzo_scene_scores(const uint64_t *sad, int n_frames, int width, int height,
                       int bitdepth, double threshold,
                       double *score, uint8_t *selected)
{
    double prev_mafd = 0.0;
    uint64_t count = (uint64_t)width * (uint64_t)height;
    for (int t = 0; t < n_frames; t++) {
        double ret = 0.0;
        if (t > 0) {
            double mafd = (double)sad[t] / count / (1ULL << (bitdepth - 8));
            double diff = fabs(mafd - prev_mafd);
            double m = mafd > diff ? diff : mafd; /* FFMIN(mafd, diff) */
            ret = clipf((float)(m / 100.), 0, 1); /* double -> float32 -> double */
            prev_mafd = mafd;
        }
        if (score) score[t] = ret;
        if (selected) selected[t] = ret > threshold ? 1 : 0;
    }
}

/* One stream, frame by frame, exactly as the filter sees it: SAD against the
 * previous frame, then the score.  sad_out[0] = 0. */
void tvzo_scene_stream(const uint8_t *luma, int n_frames, int width, int height,
                       int64_t pitch, int64_t frame_stride, int bitdepth,
                       double threshold, uint64_t *sad_out, double *score,
                       uint8_t *selected)
{
    if (n_frames <= 0) return;
    sad_out[0] = 0;
    for (int t = 1; t < n_frames; t++)
        sad_out[t] = (bitdepth > 8 ? tvzo_scene_sad_u16 : tvzo_scene_sad_u8)(
            luma + (int64_t)(t - 1) * frame_stride, pitch, luma + (int64_t)t * frame_stride, pitch, width, height);
    tvzo_scene_scores(sad_out, n_frames, width, height, bitdepth, threshold,
                      score, selected);
}

/* Many independent streams (the reference runs one ffmpeg process per upload,
 * app.py:43,472): one stream per OpenMP thread.  n_threads <= 0 -> all cores.
 * Returns the number of threads used. */
int tvzo_scene_batch(const uint8_t *luma, int n_streams, int n_frames, int width,
                     int height, int64_t pitch, int64_t frame_stride,
                     int64_t stream_stride, int bitdepth, double threshold,
                     uint64_t *sad_out, double *score, uint8_t *selected,
                     int n_threads)
{
    int used = 1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    if (n_threads > n_streams) n_threads = n_streams > 0 ? n_streams : 1;
    used = n_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int s = 0; s < n_streams; s++) {
        tvzo_scene_stream(luma + (int64_t)s * stream_stride, n_frames, width,
                          height, pitch, frame_stride, bitdepth, threshold,
                          sad_out + (int64_t)s * n_frames,
                          score ? score + (int64_t)s * n_frames : NULL,
                          selected ? selected + (int64_t)s * n_frames : NULL);
    }
    return used;
}

/* av_ts_make_time_string (libavutil/timestamp.h, FFmpeg <= 6.x):
 *   snprintf(buf, 32, "%.6g", av_q2d(tb) * ts),  av_q2d = num / (double)den.
 * mode 0 = "%.6g" (FFmpeg 4.x-6.x); mode 1 = FFmpeg >= 7.0
 * av_ts_make_time_string2: "%.*f" with precision 6 (more when |val| < 1) and
 * trailing zeros / a trailing '.' trimmed.  vf_showinfo prints this after the
 * token `pts_time:`, which the reference parses at app.py:228-230. */
int tvzo_pts_time_string(int64_t pts, int tb_num, int tb_den, int mode,
                         char *buf, int buflen)
{
    double val = ((double)tb_num / (double)tb_den) * (double)pts;
    if (mode == 0)
        return snprintf(buf, (size_t)buflen, "%.6g", val);
    {
        double lg = floor(log10(fabs(val)));
        int precision = (isfinite(lg) && lg < 0) ? (int)(-lg) + 5 : 6;
        int last = snprintf(buf, (size_t)buflen, "%.*f", precision, val);
        if (last > buflen - 1) last = buflen - 1;
        last -= 1;
        for (; last && buf[last] == '0'; last--) ;
        for (; last && buf[last] != 'f' && (buf[last] < '0' || buf[last] > '9'); last--) ;
        buf[last + 1] = '\0';
        return last + 1;
    }
}

/* The parse at app.py:228-232: float(token) of every selected frame, appended
 * iff it differs from the last appended value.  Returns the number of cuts. */
int tvzo_cut_timestamps(const uint8_t *selected, int n_frames, const int64_t *pts,
                        int tb_num, int tb_den, int mode, double *out)
{
    int n = 0;
    char buf[64];
    for (int t = 0; t < n_frames; t++) {
        if (!selected[t]) continue;
        tvzo_pts_time_string(pts ? pts[t] : (int64_t)t, tb_num, tb_den, mode, buf, sizeof buf);
        double ts = strtod(buf, NULL);
        if (n == 0 || ts != out[n - 1]) out[n++] = ts;
    }
    return n;
}
