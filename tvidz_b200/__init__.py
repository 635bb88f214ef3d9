"""tvidz_b200 -- B200-native implementation of the TVIDZ analysis hot path.

Two stages, both hand-written sm_100a CUDA behind the C ABI of include/tvidz_b200.h:
  scene   : FFmpeg `select='gt(scene,T)'` scoring (luma byte-SAD -> mafd -> score -> cuts)
  catalog : find_duplicates of a cut-timestamp list against the packed catalogue
`inspector` mirrors the reference's function names and return formats for this path.
There is no CPU fallback: importing the package is cheap, but every compute entry point
loads tvidz_b200/libtvidz_b200.so and raises if it is missing.
"""
from . import _lib, catalog, fragment, inspector, scene, synth  # noqa: F401
from .catalog import Catalogue  # noqa: F401
from .fragment import FragmentCatalogue  # noqa: F401
from .inspector import Inspector  # noqa: F401
from .scene import detect_scene_cuts, score_frames, score_frames_host  # noqa: F401

__all__ = ["Catalogue", "Inspector", "detect_scene_cuts", "score_frames", "score_frames_host",
           "scene", "catalog", "inspector", "synth"]
