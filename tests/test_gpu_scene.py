"""Stage 1 parity on the GPU: the CUDA path (through the C ABI) against the CPU oracle on
the same seeded inputs -- bit-exact integer SADs, scores, selections and cut lists."""
import numpy as np
import pytest
import torch

import oracle
import tvidz_b200._lib as L
from tvidz_b200 import scene, synth

pytestmark = pytest.mark.gpu


def _check(frames_cpu: torch.Tensor, dev, width=None, threshold=0.3, expect_path=None):
    f = frames_cpu.numpy()
    S, F, H, P = f.shape
    W = P if width is None else width
    o_sad, o_score, o_sel, _ = oracle.scene_batch(np.ascontiguousarray(f), W, threshold)
    d = frames_cpu.to(dev)
    if expect_path is not None:
        pitch, fs, ss = scene._strides(d)
        assert L.lib().tvz_sad_luma_u8_path(d.data_ptr(), W, H, pitch, fs, ss) == expect_path
    sad, score, sel = scene.score_frames(d, W, threshold)
    torch.cuda.synchronize()
    assert np.array_equal(sad.cpu().numpy().astype(np.uint64), o_sad)
    assert np.array_equal(score.cpu().numpy(), o_score)           # bit-exact doubles
    assert np.array_equal(sel.cpu().numpy(), o_sel)
    return o_sad, o_sel


@pytest.mark.parametrize("S,F,H,W", [(1, 2, 16, 16), (3, 9, 48, 64), (2, 33, 270, 480), (5, 7, 1, 4096),
                                     (1, 5, 1080, 1920), (4, 18, 720, 1280), (2, 3, 2160, 3840)])
def test_flat_layout_random(cuda, S, F, H, W):
    g = torch.Generator().manual_seed(S * 1000 + F)
    frames = torch.randint(0, 256, (S, F, H, W), dtype=torch.uint8, generator=g)
    _check(frames, cuda, expect_path=1)


@pytest.mark.parametrize("S,F,H,W,P", [(2, 6, 33, 64, 128), (1, 4, 1080, 1920, 2048), (3, 5, 17, 4096, 4112),
                                       (2, 9, 2160, 3840, 4096)])
def test_row_band_layout_pitch_gt_width(cuda, S, F, H, W, P):
    g = torch.Generator().manual_seed(P)
    frames = torch.randint(0, 256, (S, F, H, P), dtype=torch.uint8, generator=g)   # random padding too
    _check(frames, cuda, width=W, expect_path=1)


@pytest.mark.parametrize("S,F,H,W,P", [(2, 5, 31, 1918, 1984), (1, 4, 9, 13, 13), (3, 3, 20, 50, 51),
                                       (1, 6, 270, 479, 479), (2, 4, 5, 1, 7)])
def test_generic_layouts(cuda, S, F, H, W, P):
    g = torch.Generator().manual_seed(W)
    frames = torch.randint(0, 256, (S, F, H, P), dtype=torch.uint8, generator=g)
    _check(frames, cuda, width=W, expect_path=0)


def test_unaligned_base_takes_generic_path(cuda):
    g = torch.Generator().manual_seed(1)
    big = torch.randint(0, 256, (2 * 6 * 32 * 64 + 16,), dtype=torch.uint8, generator=g)
    view_cpu = big[3:3 + 2 * 6 * 32 * 64].view(2, 6, 32, 64)
    d = big.to(cuda)[3:3 + 2 * 6 * 32 * 64].view(2, 6, 32, 64)
    assert L.lib().tvz_sad_luma_u8_path(d.data_ptr(), 64, 32, 64, 32 * 64, 6 * 32 * 64) == 0
    sad = scene.sad_luma(d)
    o_sad, _, _, _ = oracle.scene_batch(np.ascontiguousarray(view_cpu.numpy()))
    assert np.array_equal(sad.cpu().numpy().astype(np.uint64), o_sad)


def test_known_answers_1080p(cuda):
    """SURVEY.md A.5 #1-#4 at full 1920x1080 through the CUDA path."""
    H, W = 1080, 1920
    mk = lambda vals: torch.stack([torch.full((H, W), v, dtype=torch.uint8) for v in vals])[None]
    sad, score, sel = scene.score_frames(mk([77, 77, 77]).to(cuda))
    assert sad.tolist() == [[0, 0, 0]] and sel.tolist() == [[0, 0, 0]]
    sad, score, sel = scene.score_frames(mk([0, 255, 255]).to(cuda))
    assert sad.tolist() == [[0, 528768000, 0]] and score.tolist() == [[0.0, 1.0, 0.0]] and sel.tolist() == [[0, 1, 0]]
    sad, score, sel = scene.score_frames(mk([0, 30]).to(cuda))
    assert sad[0, 1].item() == 62208000 and score[0, 1].item() == float(np.float32(0.3)) and sel[0, 1].item() == 1
    f = mk([0, 30])
    f[0, 1, 0, 0] = 29                                      # sad = 62,207,999 -> float32 0.29999998 -> no cut
    sad, score, sel = scene.score_frames(f.to(cuda))
    assert sad[0, 1].item() == 62207999 and score[0, 1].item() < 0.3 and sel[0, 1].item() == 0
    sad, score, sel = scene.score_frames(mk([0, 40, 80, 120, 160, 200]).to(cuda))
    assert sel.tolist() == [[0, 1, 0, 0, 0, 0]] and score[0, 1].item() == float(np.float32(0.4))


def test_max_sad_4k_fits(cuda):
    H, W = 2160, 3840
    f = torch.zeros((1, 3, H, W), dtype=torch.uint8)
    f[0, 1] = 255
    sad = scene.sad_luma(f.to(cuda))
    assert sad.tolist() == [[0, 2115072000, 2115072000]]


def test_empty_and_single_frame(cuda):
    f = torch.zeros((2, 1, 8, 16), dtype=torch.uint8, device=cuda)
    sad, score, sel = scene.score_frames(f)
    assert sad.tolist() == [[0], [0]] and score.tolist() == [[0.0], [0.0]] and sel.tolist() == [[0], [0]]
    assert scene.sad_luma(torch.zeros((0, 4, 8, 16), dtype=torch.uint8, device=cuda)).shape == (0, 4)


@pytest.mark.parametrize("variant,ctas,units,minseg", [(0, 0, 16, 16), (1, 0, 1, 1), (2, 2, 64, 2), (3, 1, 4, 4),
                                                       (4, 0, 16, 3), (5, 1, 200, 1)])
def test_ring_variants_and_time_segments(cuda, variant, ctas, units, minseg):
    """Every ring geometry and aggressive time segmentation give the same integers."""
    frames = synth.synth_frames(3, 41, 90, 160, seed=variant, scene_len=(4, 11))
    try:
        L.check(L.lib().tvz_debug_sad_tuning(variant, ctas, units, minseg))
        _, o_sel = _check(frames, cuda, expect_path=1)
        assert o_sel.sum() >= 6                                  # the synthetic cuts are really there
    finally:
        L.check(L.lib().tvz_debug_sad_tuning(0, 0, 16, 16))


def test_synthetic_streams_cut_lists(cuda):
    """Synthetic scenes: exactly one cut per scene change, cut lists identical
    to the oracle's, device-resident and host-buffer entries alike."""
    frames = synth.synth_frames(4, 64, 135, 240, seed=7, scene_len=(5, 20))
    f = frames.numpy()
    _, _, o_sel, _ = oracle.scene_batch(f)
    want = [oracle.cut_timestamps(o_sel[s]) for s in range(4)]
    assert all(len(w) >= 2 for w in want)
    assert scene.detect_scene_cuts(frames.to(cuda)) == want
    assert scene.detect_scene_cuts(frames.pin_memory()) == want          # host-buffer entry (torch)
    assert scene.detect_scene_cuts(f) == want                            # host-buffer entry (numpy)
    assert scene.detect_scene_cuts(frames[0].to(cuda)) == want[0]


@pytest.mark.parametrize("chunk", [1, 2, 5, 7, 64, 0])
def test_host_entry_chunking(cuda, chunk):
    frames = synth.synth_frames(3, 23, 54, 96, seed=chunk, scene_len=(3, 9)).pin_memory()
    o_sad, o_score, o_sel, _ = oracle.scene_batch(frames.numpy())
    sad, score, sel = scene.score_frames_host(frames, chunk_frames=chunk)
    assert np.array_equal(sad, o_sad) and np.array_equal(score, o_score) and np.array_equal(sel, o_sel)


def test_host_entry_padded_rows(cuda):
    frames = synth.synth_frames(2, 12, 30, 50, seed=3, scene_len=(3, 6), pitch=64)
    o_sad, o_score, o_sel, _ = oracle.scene_batch(frames.numpy(), 50)
    sad, score, sel = scene.score_frames_host(frames.numpy(), width=50, chunk_frames=5)
    assert np.array_equal(sad, o_sad) and np.array_equal(sel, o_sel)


def test_full_size_properties(cuda):
    """BASELINE config 2 geometry (64 streams of 1080p; fewer frames): size-independent checks
    -- time reversal leaves each SAD in place mirrored, a constant offset on both frames
    changes nothing, and a sampled stream equals the oracle."""
    S, F, H, W = 64, 6, 1080, 1920
    g = torch.Generator(device=cuda).manual_seed(5)
    frames = torch.randint(16, 236, (S, F, H, W), dtype=torch.uint8, device=cuda, generator=g)
    sad = scene.sad_luma(frames)
    rev = scene.sad_luma(frames.flip(1).contiguous())
    assert torch.equal(sad[:, 1:], rev[:, 1:].flip(1))
    assert torch.equal(scene.sad_luma(frames + 7), sad)
    for s in (0, 37, 63):
        o_sad, _, _, _ = oracle.scene_batch(frames[s:s + 1].cpu().numpy())
        assert np.array_equal(sad[s:s + 1].cpu().numpy().astype(np.uint64), o_sad)


@pytest.mark.parametrize("chunks", [[1, 1, 10], [5, 7], [12]])
def test_stream_scorer_ten_bit_chunks(cuda, chunks):
    """16-bit samples fed chunk by chunk: the carried frame keeps its 16 bits and mafd is divided by
    2^(bitdepth-8) for the carry pair as for the chunk -- identical to one call and to the oracle; the
    P016 layout NVDEC writes (sample in the HIGH bits) scores the same with bitdepth = 16."""
    rng = np.random.default_rng(sum(chunks))
    base = rng.integers(100, 600, (2, 1, 40, 64))
    f = base + rng.integers(0, 8, (2, 12, 40, 64))                                          # one scene, small noise ...
    f[:, 5:] += 300                                                                          # ... and a cut at frame 5
    f = f.astype(np.uint16)
    o_sad, o_score, o_sel, _ = oracle.scene_batch(f, bitdepth=10)
    for shift, depth in ((0, 10), (6, 16)):
        d = torch.from_numpy((f << shift).view(np.int16)).to(cuda)
        sc = scene.StreamScorer(bitdepth=depth)
        got, t = [], 0
        for n in chunks:
            got.append(sc.feed(d[:, t:t + n]))
            t += n
        sad = torch.cat([g[0] for g in got], 1).cpu().numpy().astype(np.uint64)
        assert np.array_equal(sad, o_sad << np.uint64(shift))
        assert np.array_equal(torch.cat([g[1] for g in got], 1).cpu().numpy(), o_score)
        assert np.array_equal(torch.cat([g[2] for g in got], 1).cpu().numpy(), o_sel) and o_sel[:, 5].all()
    with pytest.raises((TypeError, ValueError)):
        scene.StreamScorer(bitdepth=8).feed(torch.from_numpy(f.view(np.int16)).to(cuda))


@pytest.mark.parametrize("chunks", [[1, 1, 1, 37], [16, 16, 8], [40], [7, 0, 33], [39, 1]])
def test_stream_scorer_equals_one_shot(cuda, chunks):
    """Long-form video fed chunk by chunk (BASELINE config 3 shape: 4K frames) == one call."""
    H, W = 2160, 3840
    frames = synth.synth_frames(1, 40, H, W, seed=9, scene_len=(6, 13)).to(cuda)
    sad, score, sel = scene.score_frames(frames)
    sc = scene.StreamScorer()
    got, t = [], 0
    for n in chunks:
        got.append(sc.feed(frames[0, t:t + n]))
        t += n
    assert t == 40 and sc.frames_seen == 40
    assert torch.equal(torch.cat([g[0] for g in got], 1), sad)
    assert torch.equal(torch.cat([g[1] for g in got], 1), score)
    assert torch.equal(torch.cat([g[2] for g in got], 1), sel)
    assert int(sel.sum()) >= 3
    o_sad, o_score, o_sel, _ = oracle.scene_batch(frames.cpu().numpy())
    assert np.array_equal(sad.cpu().numpy().astype(np.uint64), o_sad) and np.array_equal(sel.cpu().numpy(), o_sel)


def test_config1_single_clip_end_to_end(cuda):
    """BASELINE config 1: one 60 s 1080p30 clip through scene-cut detection (T = 0.3) and the compare
    against 10 stored timestamp arrays -- the whole analyze_file path vs the oracle's."""
    from oracle import match_oracle
    from tvidz_b200.inspector import Inspector
    frames = synth.synth_frames(1, 1800, 1080, 1920, seed=60, scene_len=(45, 240), device=cuda)
    o_sad, o_score, o_sel, _ = oracle.scene_batch(frames.cpu().numpy())
    want_cuts = oracle.cut_timestamps(o_sel[0])
    assert 6 <= len(want_cuts) <= 40
    assert scene.detect_scene_cuts(frames[0]) == want_cuts
    rng = np.random.default_rng(60)
    stored = [("v%d.mp4" % i, sorted(float("%.6g" % (n / 30)) for n in rng.integers(1, 1800, 12))) for i in range(9)]
    stored.insert(4, ("original.mp4", list(want_cuts)))                 # the clip itself was uploaded before
    ins = Inspector()
    rows, names = [], {}
    for fn, ts in stored:
        v = ins.add_video(fn)
        ins.add_timestamps(v.id, ts)
        rows.append((v.id, ts))
        names[v.id] = fn
    res = ins.analyze_frames("uploads/1760000000000-original.mp4", frames[0])
    o_scene, o_ids, o_names = match_oracle.streaming_analysis(rows, 11, want_cuts, 2, names)
    assert res == {**match_oracle.result_record(o_scene, o_names, "1760000000000-original.mp4", "original.mp4"),
                   "duplicates": res["duplicates"]}
    assert set(res["duplicates"]) == set(o_names) == {"original.mp4"} and res["total_cuts"] == 2
    fresh = ins.analyze_frames("other.mp4", synth.synth_frames(1, 300, 1080, 1920, seed=61, scene_len=(45, 90),
                                                               device=cuda)[0])
    assert fresh["status"] == "done" and fresh["duplicates"] == [] and fresh["total_cuts"] >= 2


def test_random_shapes_and_strides(cuda):
    """Hypothesis-style sweep: 48 seeded random geometries (odd widths, padded pitches, frame and
    stream strides with gaps, aligned and unaligned bases) -- every path must give the oracle's integers."""
    rng = np.random.default_rng(2026)
    paths = set()
    for case in range(48):
        S, F = int(rng.integers(1, 5)), int(rng.integers(1, 12))
        H = int(rng.integers(1, 70))
        W = int(rng.choice([1, 7, 16, 31, 48, 64, 100, 128, 240, 333, 1024, 1920]))
        kind = case % 4
        P = W if kind == 0 else W + int(rng.choice([0, 1, 16, 32, 13]))
        if kind == 2:
            P = (W + 15) // 16 * 16 + 16 * int(rng.integers(0, 3))          # 16-byte aligned pitch
        fgap = int(rng.choice([0, 16, 48])) if kind != 3 else int(rng.integers(0, 9))
        sgap = int(rng.choice([0, 32, 256])) if kind != 3 else int(rng.integers(0, 9))
        base_off = 0 if kind != 3 else int(rng.integers(0, 16))
        fstride = H * P + fgap
        sstride = F * fstride + sgap
        total = base_off + S * sstride + 64
        buf = torch.from_numpy(rng.integers(0, 256, total, dtype=np.uint8))
        view = buf[base_off:].as_strided((S, F, H, P), (sstride, fstride, P, 1))
        dense = np.ascontiguousarray(view.numpy())
        o_sad, o_score, o_sel, _ = oracle.scene_batch(dense, W)
        dbuf = buf.to(cuda)
        dview = dbuf[base_off:].as_strided((S, F, H, P), (sstride, fstride, P, 1))
        paths.add(L.lib().tvz_sad_luma_u8_path(dview.data_ptr(), W, H, P, fstride, sstride))
        sad, score, sel = scene.score_frames(dview, W)
        assert np.array_equal(sad.cpu().numpy().astype(np.uint64), o_sad), (case, S, F, H, W, P, fstride, sstride)
        assert np.array_equal(score.cpu().numpy(), o_score) and np.array_equal(sel.cpu().numpy(), o_sel)
    assert paths == {0, 1}


@pytest.mark.parametrize("S,F,H,W,P", [(2, 6, 40, 64, 64), (1, 4, 1080, 1920, 1920), (3, 5, 33, 100, 104),
                                       (2, 7, 17, 63, 63), (1, 3, 2160, 3840, 3840)])
def test_ten_bit_luma(cuda, S, F, H, W, P):
    """yuv420p10: 16-bit samples through ff_scene_sad16_c's arithmetic and mafd / 4 (A.1, A.3)."""
    rng = np.random.default_rng(W + F)
    f = rng.integers(0, 1024, (S, F, H, P), dtype=np.uint16)
    f[:, F // 2:] = np.clip(f[:, F // 2:].astype(np.int32) + 300, 0, 1023).astype(np.uint16)   # a cut
    o_sad, o_score, o_sel, _ = oracle.scene_batch(f, W, bitdepth=10)
    d = torch.from_numpy(f.view(np.int16)).to(cuda)
    sad, score, sel = scene.score_frames(d, W, bitdepth=10)
    assert np.array_equal(sad.cpu().numpy().astype(np.uint64), o_sad)
    assert np.array_equal(score.cpu().numpy(), o_score) and np.array_equal(sel.cpu().numpy(), o_sel)
    h_sad, h_score, h_sel = scene.score_frames_host(f, W, bitdepth=10, chunk_frames=3)
    assert np.array_equal(h_sad, o_sad) and np.array_equal(h_score, o_score) and np.array_equal(h_sel, o_sel)
    assert o_sel.sum() >= S
    # full-scale KAT: 0 -> 1023 everywhere is mafd 1023/4 = 255.75 -> score clips to 1.0
    k = np.zeros((1, 2, 8, 16), np.uint16)
    k[0, 1] = 1023
    sad, score, sel = scene.score_frames(torch.from_numpy(k.view(np.int16)).to(cuda), bitdepth=10)
    assert sad.tolist() == [[0, 1023 * 128]] and score.tolist() == [[0.0, 1.0]]


def test_concurrent_analyses_like_the_reference(cuda):
    """The reference runs analyze_file on one thread per upload (app.py:43,472): several threads go
    through the host-buffer entry and the Inspector at once and must get what a serial run gets."""
    import threading
    from tvidz_b200.inspector import Inspector
    clips = [synth.synth_frames(1, 90, 270, 480, seed=200 + i, scene_len=(6, 15))[0] for i in range(6)]
    want = []
    for c in clips:
        _, _, o_sel, _ = oracle.scene_batch(c[None].numpy())
        want.append(oracle.cut_timestamps(o_sel[0]))
    ins = Inspector()
    for i in (0, 3):                                   # two of the clips were uploaded before
        v = ins.add_video("old-%d.mp4" % i)
        ins.add_timestamps(v.id, want[i])
    results = [None] * len(clips)

    def work(i):
        for _ in range(3):
            sad, score, sel = scene.score_frames_host(clips[i].numpy(), chunk_frames=16)
            assert scene.cut_timestamps(sel[0]) == want[i]
        results[i] = ins.analyze_frames("up-%d.mp4" % i, clips[i].numpy())

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(clips))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i, r in enumerate(results):
        assert r is not None and r["status"] == "done", (i, r)
        if i in (0, 3):
            assert r["duplicates"] == ["old-%d.mp4" % i] and r["scene_cuts"] == want[i][:2]
        else:
            assert r["scene_cuts"] == want[i] or r["duplicates"]      # may match an earlier concurrent upload only by chance
            assert r["total_cuts"] == len(r["scene_cuts"])


@pytest.mark.parametrize("H,W,P", [(270, 480, 480 * 3), (1080, 1920, 1920 * 3), (33, 101, 320)])
def test_packed_rgb24_is_a_3w_byte_plane(cuda, H, W, P):
    """Packed RGB24 / BGR24 sources (FFmpeg's select filter keeps them as ONE plane whose
    visible width is av_image_get_linesize = 3*w bytes, count = 3*w*h): the same SAD and score kernels
    with width = 3 * w -- no separate entry point is needed."""
    g = torch.Generator().manual_seed(H + W)
    frames = torch.empty((2, 6, H, P), dtype=torch.uint8)
    frames[:, :3] = torch.randint(0, 256, (2, 1, H, P), dtype=torch.uint8, generator=g)   # three stills,
    frames[:, 3:] = torch.randint(0, 256, (2, 1, H, P), dtype=torch.uint8, generator=g)   # a hard cut, three stills
    o_sad, o_sel = _check(frames, cuda, width=3 * W)
    assert o_sel[:, 3].all() and not o_sel[:, 4:].any()
