"""Run the UNMODIFIED reference modules (/root/reference/inspector/db.py and
app.py) in this container with their I/O dependencies stubbed, so that the
reference's own `find_duplicates` and `analyze_file` code paths produce the
golden vectors committed next to this file.

Only usable where /root/reference exists (the build container); the GPU box
never runs this -- tests read the committed JSON fixtures instead.

What is stubbed (none of it is on the arithmetic path):
  sqlalchemy            -> an in-memory session; rows are copied through an
                           8-byte pack/unpack on read so list elements are new
                           float objects, as after a psycopg2 float8[] round trip
  flask, boto3, ffmpeg  -> inert modules (routes are never exercised)
  requests.get          -> returns a few bytes for the "download"
  subprocess.Popen      -> a fake ffmpeg whose stderr yields showinfo lines
"""
from __future__ import annotations

import importlib
import struct
import sys
import types

REF_DIR = "/root/reference/inspector"


def _fresh_float(x):
    return struct.unpack("<d", struct.pack("<d", float(x)))[0]


class _Store:
    def __init__(self):
        self.rows = {}
        self.next_id = {}

    def clear(self):
        self.rows.clear()
        self.next_id.clear()


STORE = _Store()


class _Query:
    def __init__(self, model, filt=None):
        self.model, self.filt = model, filt or {}

    def filter_by(self, **kw):
        f = dict(self.filt)
        f.update(kw)
        return _Query(self.model, f)

    def _matching(self):
        out = []
        for r in STORE.rows.get(self.model.__name__, []):
            if all(getattr(r, k, None) == v for k, v in self.filt.items()):
                out.append(r)
        return out

    def _hydrate(self, r):
        # what the ORM hands back for a float8[] column: a new list of new floats
        if hasattr(r, "timestamps") and isinstance(r.timestamps, list):
            clone = types.SimpleNamespace(**r.__dict__)
            clone.timestamps = [_fresh_float(x) for x in r.timestamps]
            clone._orig = r
            return clone
        return r

    def all(self):
        return [self._hydrate(r) for r in self._matching()]

    def first(self):
        m = self._matching()
        return m[0] if m else None   # un-hydrated: db.add_timestamps mutates it in place

    def delete(self):
        keep = [r for r in STORE.rows.get(self.model.__name__, []) if r not in self._matching()]
        STORE.rows[self.model.__name__] = keep


class _Session:
    def query(self, model):
        return _Query(model)

    def add(self, obj):
        name = type(obj).__name__
        nid = STORE.next_id.get(name, 1)
        STORE.next_id[name] = nid + 1
        obj.id = nid
        STORE.rows.setdefault(name, []).append(obj)

    def commit(self):
        pass

    def refresh(self, obj):
        pass

    def close(self):
        pass

    def rollback(self):
        pass


def _install_stubs():
    sa = types.ModuleType("sqlalchemy")
    sa.create_engine = lambda *a, **k: object()
    for name in ("Column", "ForeignKey"):
        setattr(sa, name, lambda *a, **k: None)
    for name in ("Integer", "String", "Float", "DateTime", "ARRAY", "Text"):
        setattr(sa, name, type(name, (), {"__init__": lambda self, *a, **k: None}))

    orm = types.ModuleType("sqlalchemy.orm")

    def declarative_base():
        class Base:
            metadata = types.SimpleNamespace(create_all=lambda *a, **k: None,
                                             drop_all=lambda *a, **k: None)

            def __init__(self, **kw):
                for k, v in kw.items():
                    setattr(self, k, v)
        return Base

    orm.declarative_base = declarative_base
    orm.sessionmaker = lambda **k: _Session
    orm.relationship = lambda *a, **k: None
    dialects = types.ModuleType("sqlalchemy.dialects")
    pg = types.ModuleType("sqlalchemy.dialects.postgresql")
    pg.ARRAY = sa.ARRAY
    sys.modules.update({"sqlalchemy": sa, "sqlalchemy.orm": orm,
                        "sqlalchemy.dialects": dialects, "sqlalchemy.dialects.postgresql": pg})

    fl = types.ModuleType("flask")

    class Flask:
        def __init__(self, *a, **k):
            pass

        def route(self, *a, **k):
            return lambda f: f

        def after_request(self, f):
            return f

    fl.Flask = Flask
    fl.request = types.SimpleNamespace()
    fl.jsonify = lambda *a, **k: (a, k)
    fl.Response = type("Response", (), {"__init__": lambda self, *a, **k: None, "headers": {}})
    sys.modules["flask"] = fl
    sys.modules["boto3"] = types.ModuleType("boto3")
    ff = types.ModuleType("ffmpeg")
    ff.probe = lambda path: {"streams": [{"codec_type": "video", "nb_frames": "1800"}]}
    sys.modules["ffmpeg"] = ff


_loaded = None


def load_reference():
    """-> (db module, app module), the reference's own, with stubs installed."""
    global _loaded
    if _loaded is None:
        _install_stubs()
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        db = importlib.import_module("db")
        app = importlib.import_module("app")
        _loaded = (db, app)
    return _loaded


def ref_find_duplicates(catalogue, query, min_match=None):
    """catalogue: [(filename, timestamps)] inserted in order through the
    reference's add_video/add_timestamps; returns (ids, result)."""
    db, _ = load_reference()
    STORE.clear()
    ids = []
    for fname, ts in catalogue:
        v = db.add_video(fname)
        db.add_timestamps(v.id, list(ts))
        ids.append(v.id)
    if min_match is None:
        return ids, db.find_duplicates(list(query))
    return ids, db.find_duplicates(list(query), min_match=min_match)


class _FakeProc:
    def __init__(self, lines):
        self.stderr = iter(lines)
        self.terminated = False

    def terminate(self):
        self.terminated = True

    def wait(self):
        return 0


def ref_analyze_file(catalogue, upload_name, stderr_lines):
    """Run the reference's analyze_file on a fake upload whose ffmpeg stderr is
    `stderr_lines`, against a catalogue inserted beforehand.  Returns
    (result_dict, video_id_of_upload, stored_duplicates, catalogue_ids, terminated)."""
    import subprocess

    import requests

    db, app = load_reference()
    STORE.clear()
    ids = []
    for fname, ts in catalogue:
        v = db.add_video(fname)
        db.add_timestamps(v.id, list(ts))
        ids.append(v.id)
    app.analysis_results.clear()
    proc = _FakeProc([ln + "\n" for ln in stderr_lines])

    class _Resp:
        def iter_content(self, chunk_size=8192):
            yield b"not a real video"

    real_popen, real_get = subprocess.Popen, requests.get
    subprocess.Popen = lambda *a, **k: proc
    requests.get = lambda *a, **k: _Resp()
    try:
        app.analyze_file("videos", upload_name)
    finally:
        subprocess.Popen, requests.get = real_popen, real_get
    assert len(app.analysis_results) == 1
    result = dict(next(iter(app.analysis_results.values())))
    me = [r for r in STORE.rows["Video"] if r.id not in ids][0]
    return result, me.id, list(getattr(me, "duplicates", None) or []), ids, proc.terminated
