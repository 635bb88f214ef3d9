"""oracle/match_oracle.py -- pure-Python restatement of the reference's stage 2.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs as the checker or the CPU arm;
never by anything under tvidz_b200/.

Follows /root/reference:
  * inspector/db.py:85-91   -- the membership count of find_duplicates
  * inspector/app.py:228-255 -- the per-cut streaming loop with early exit
  * inspector/app.py:293-302 -- the final result record

The catalogue stands in for `session.query(VideoTimestamps).all()` (db.py:83):
an ordered sequence of ``(video_id, timestamps)`` pairs, timestamps being plain
Python lists of floats so that ``in`` has exactly the reference's meaning
(list.__contains__: identity, then ``==``).

Pinned by tests/golden/match_*.json, produced by running the reference's own
find_duplicates (stubbed SQLAlchemy session) -- see tests/golden/gen_match_golden.py.
"""
from __future__ import annotations

import struct


def hydrate(rows):
    """What `session.query(VideoTimestamps).all()` (db.py:83) hands back: every row a new
    list of NEW float objects (float8[] travels as 8-byte binary).  Matters only for NaN,
    which `in` would otherwise match by object identity."""
    return [(vid, [struct.unpack("<d", struct.pack("<d", float(x)))[0] for x in ts]) for vid, ts in rows]


def find_duplicates(catalogue, new_timestamps, min_match=5):
    """db.py:76-94 over an in-memory catalogue; result in catalogue order."""
    out = []
    for video_id, stored in catalogue:
        hits = 0
        for ts in new_timestamps:       # counted over QUERY positions (B.1)
            if ts in stored:
                hits += 1
        if hits >= min_match:
            out.append((video_id, hits))
    return out


def streaming_analysis(catalogue, self_video_id, pts_time_tokens, min_match=2,
                       names=None):
    """The loop of app.py:216-255 fed with the `pts_time:` tokens of the
    selected frames, against a snapshot catalogue (SURVEY.md B.3 caveat).

    Returns (scene_timestamps, duplicate_ids, duplicate_names).  The querying
    video's own row (upserted at app.py:234) is dropped at app.py:237, so it
    need not be present in `catalogue`; if it is, it is ignored the same way.
    """
    scene = []
    dup_ids, dup_names = [], []
    for token in pts_time_tokens:
        ts = float(token)                                   # app.py:230
        if scene and ts == scene[-1]:                       # app.py:231
            continue
        scene.append(ts)
        dups = find_duplicates(catalogue, scene, min_match)  # app.py:235
        dups = [d for d in dups if d[0] != self_video_id]    # app.py:237
        if dups:
            dup_ids = [d[0] for d in dups]                   # app.py:239
            if names is not None:                            # app.py:241-245
                dup_names = [names[i] for i in dup_ids if i in names]
            break                                            # app.py:251-255
    return scene, dup_ids, dup_names


def result_record(scene, dup_names, filename, clean_filename):
    """app.py:293-302."""
    return {
        'status': 'done',
        'scene_cuts': scene,
        'progress': 1.0,
        'total_cuts': len(scene),
        'duplicates': list(set(dup_names)) if dup_names else [],
        'original_filename': filename,
        'clean_filename': clean_filename,
    }
