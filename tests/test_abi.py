"""The C-ABI library loads and exports every symbol include/tvidz_b200.h declares
(no compute calls here: this file runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import tvidz_b200._lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "tvidz_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(tvz_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(L.LIB_PATH)
    bound = {n for n, _, _ in L.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in bound, f"{n} declared in the header but not bound in _lib.py"
    assert bound <= set(names), "binding names a symbol the header does not declare"


def test_abi_version_and_error_string():
    lib = L.lib()
    assert lib.tvz_abi_version() == 1
    assert isinstance(lib.tvz_last_error(), bytes)


def test_argument_validation_needs_no_gpu():
    lib = L.lib()
    rc = lib.tvz_sad_luma_u8(None, 1, 2, 16, 16, 16, 256, 512, None, None)
    assert rc == -1 and b"null" in lib.tvz_last_error()
    rc = lib.tvz_sad_luma_u8(1, 1, 2, 16, 16, 8, 256, 512, 1, None)
    assert rc == -1 and b"pitch" in lib.tvz_last_error()
    with pytest.raises(L.TvzError):
        L.check(rc)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libtvidz_b200.so")
    with pytest.raises(RuntimeError, match="no CPU path"):
        L.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tvidz_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "oracle/" not in src and "tvzo_" not in src, fn


def test_fingerprint_layout_is_a_conflict_free_permutation():
    """Host logic of the catalogue packer (no GPU): inside every 512-value unit the arranged
    fingerprints are a permutation of the values' hashes, and the k-th lookups of the 32 lanes
    (positions lane * 16 + k) fall into nearly 32 different shared-memory banks (banks that hold
    more than 16 of a unit's values must double up somewhere)."""
    import numpy as np
    from tvidz_b200 import _lib, synth
    ts, off, vid = synth.synth_catalogue(400, seed=5)
    n = int(ts.shape[0])
    padded = -(-n // 512) * 512
    fp = np.zeros(padded, np.uint16)
    perm = np.zeros(padded, np.uint16)
    _lib.check(_lib.lib().tvz_debug_arrange_fingerprints(ts.ctypes.data, n, fp.ctypes.data, perm.ctypes.data))
    bits = ts.view(np.uint64)
    lo, hi = (bits & np.uint64(0xffffffff)).astype(np.uint64), (bits >> np.uint64(32)).astype(np.uint64)
    h = (((lo * np.uint64(0x9E3779B1) + hi * np.uint64(0x85EBCA77)) & np.uint64(0xffffffff)) >> np.uint64(16)).astype(np.uint16)
    worst = []
    for u in range(padded // 512):
        p = perm[u * 512:(u + 1) * 512].astype(np.int64)
        assert sorted(p.tolist()) == list(range(512))                       # a permutation of the unit
        idx = u * 512 + p
        want = np.where(idx < n, h[np.minimum(idx, n - 1)], 0)
        assert np.array_equal(fp[u * 512:(u + 1) * 512], want)             # every position carries its value's hash
        if (u + 1) * 512 <= n:
            banks = ((fp[u * 512:(u + 1) * 512] >> 2) & 31).reshape(32, 16)  # [lane][round]
            worst.append(np.mean([np.bincount(banks[:, k], minlength=32).max() for k in range(16)]))
    assert worst and np.mean(worst) < 1.75, np.mean(worst)                  # ~1.58; random placement gives ~3.6


def test_fingerprint_layout_edge_sizes():
    """Units that are empty, partial, exactly full, or dominated by one bank (equal values, padding)
    still come out as permutations that carry the right hashes."""
    import numpy as np
    from tvidz_b200 import _lib
    rng = np.random.default_rng(11)
    for n, kind in ((0, "rand"), (1, "rand"), (511, "rand"), (512, "rand"), (513, "rand"), (2048, "same"), (1500, "few")):
        if kind == "same":
            ts = np.full(n, 12.5)
        elif kind == "few":
            ts = rng.choice(np.array([1.0, 2.5, 1e9, 0.0]), n)
        else:
            ts = np.round(rng.uniform(0, 7200, n) * 30) / 30
        ts = np.ascontiguousarray(ts, np.float64)
        padded = max(1, -(-n // 512)) * 512
        fp = np.full(padded, 0xffff, np.uint16)
        perm = np.full(padded, 0xffff, np.uint16)
        _lib.check(_lib.lib().tvz_debug_arrange_fingerprints(ts.ctypes.data if n else None, n, fp.ctypes.data,
                                                             perm.ctypes.data))
        bits = ts.view(np.uint64)
        lo, hi = bits & np.uint64(0xffffffff), bits >> np.uint64(32)
        h = (((lo * np.uint64(0x9E3779B1) + hi * np.uint64(0x85EBCA77)) & np.uint64(0xffffffff)) >> np.uint64(16)).astype(np.uint16)
        for u in range(padded // 512):
            p = perm[u * 512:(u + 1) * 512].astype(np.int64)
            assert sorted(p.tolist()) == list(range(512)), (n, kind, u)
            idx = u * 512 + p
            want = np.where(idx < n, h[np.minimum(idx, max(n - 1, 0))] if n else 0, 0)
            assert np.array_equal(fp[u * 512:(u + 1) * 512], want), (n, kind, u)


def _tiles(off, want):
    import ctypes as C
    lib = L.lib()
    off = np.ascontiguousarray(off, np.int64)
    out = np.zeros((20000, 4), np.int32)
    U = C.c_int32(-1)
    n = lib.tvz_debug_build_tiles(off.ctypes.data, off.shape[0] - 1, want, out.ctypes.data, out.shape[0], C.byref(U))
    assert 0 <= n <= out.shape[0]
    return out[:n], int(U.value)


@pytest.mark.parametrize("shape", ["cfg4", "short_rows", "empty_rows", "giant_row", "tiny", "all_empty"])
def test_tiling_covers_every_row_once(shape):
    """The catalogue cut the matcher's CTAs work on (host only): whole rows, every row in exactly one tile,
    at most 4096 rows per tile, and a unit range that holds every value of the tile's rows.  Regular
    catalogues get the arithmetic ("uniform") cut, awkward ones the greedy one."""
    rng = np.random.default_rng(3)
    if shape == "cfg4":
        lens = rng.integers(8, 121, 300_000)
    elif shape == "short_rows":
        lens = rng.integers(0, 3, 2_000_000)
    elif shape == "empty_rows":
        lens = np.zeros(30_000, np.int64)
        lens[rng.integers(0, 30_000, 500)] = rng.integers(1, 90, 500)
    elif shape == "giant_row":
        lens = rng.integers(1, 30, 3000)
        lens[1500] = 400_000
    elif shape == "tiny":
        lens = np.array([1, 1])
    else:
        lens = np.zeros(10, np.int64)
    off = np.zeros(lens.shape[0] + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    for want in (295, 296, 7):
        tiles, U = _tiles(off, want)
        assert tiles[:, 1].max() <= 4096 and tiles[:, 1].min() >= 0
        assert tiles[0, 0] == 0 and np.array_equal(tiles[1:, 0], tiles[:-1, 0] + tiles[:-1, 1])
        assert tiles[-1, 0] + tiles[-1, 1] == lens.shape[0]
        for t, (r0, nr, u0, u1) in enumerate(tiles.tolist()):
            lo, hi = int(off[r0]), int(off[r0 + nr])
            if hi > lo:
                assert u0 * 512 <= lo and hi <= u1 * 512, (shape, want, t)
            if U > 0:
                assert u0 == t * U
        if shape == "cfg4" and want > 7:
            assert U > 0 and want - 2 <= len(tiles) <= want     # one wave, balanced by stored values
            assert np.ptp(tiles[:-1, 3] - tiles[:-1, 2]) <= 2
        if shape == "empty_rows":
            assert U == 0                                        # thousands of rows at one offset: greedy cut
