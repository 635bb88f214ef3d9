"""Small end-to-end pass of the matcher (single, batched) and the fragment matcher (anchor 1..3)
against the oracle: a quick target for memory checkers / debuggers where they are available
."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from tvidz_b200 import synth
from tvidz_b200.catalog import Catalogue
from tvidz_b200.fragment import FragmentCatalogue, clip_query

ts, off, vid = synth.synth_catalogue(3000, seed=4)
cat = Catalogue(ts, off, vid, hit_capacity=64)
for r, mm in ((5, 2), (2999, 5), (100, 0)):
    q = ts[off[r]:off[r + 1]]
    assert cat.find_duplicates(q, mm) == oracle.find_duplicates_csr(ts, off, vid, q, mm)
many = cat.find_duplicates_many([ts[off[r]:off[r + 1]] for r in (1, 2, 3)], 2)
assert many[1] == oracle.find_duplicates_csr(ts, off, vid, ts[off[2]:off[3]], 2)
cat.close()
# device-side upserts: replace a packed row, append, rewrite the last row in place, then query both kernels
mcat = Catalogue(ts, off, vid, hit_capacity=64, mutable=True, tail_values=2048)
model = {int(vid[r]): ts[off[r]:off[r + 1]].tolist() for r in range(3000)}
for v, row in ((7, [1.5, 2.5]), (9001, ts[off[5]:off[6]].tolist()), (9001, ts[off[5]:off[6]].tolist() + [9.25]), (7, [])):
    assert mcat.upsert(v, row)
    model.pop(v, None)
    model[v] = row
from tvidz_b200.catalog import rows_to_csr
mts, moff, mvid = rows_to_csr(list(model.items()))
for q, mm in ((ts[off[5]:off[6]], 2), ([1.5, 2.5], 1), ([9.25], 0)):
    assert mcat.find_duplicates(q, mm) == oracle.find_duplicates_csr(mts, moff, mvid, q, mm)
assert mcat.find_duplicates_many([ts[off[5]:off[6]], [9.25]], 1) == [oracle.find_duplicates_csr(mts, moff, mvid, q, 1) for q in (ts[off[5]:off[6]], [9.25])]
mcat.close()
fts, foff, fvid = synth.synth_catalogue(300, len_range=(600, 1400), gap_range=(15, 150), seed=3)
fq = clip_query(fts[foff[21]:foff[22]], 30_000)
fcat = FragmentCatalogue(fts, foff, fvid)
for anchor in (1, 2, 3):
    v, s, d = fcat.match(fq, 4, anchor=anchor)
    assert list(zip(v.tolist(), s.tolist(), d.tolist())) == oracle.find_fragments_csr(fts, foff, fvid, fq, min_match=4, anchor=anchor)
fcat.close()
print("sanitize_small ok")
