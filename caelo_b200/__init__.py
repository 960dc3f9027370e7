"""caelo_b200 — the CAE-LO odometry hot path on NVIDIA B200 (sm_100a).

Public surface = the reference's own function names (see api.py); device work is done by the
hand-written CUDA kernels in csrc/ reached through the C ABI declared in include/caelo.h."""
__version__ = "0.1.0"

_API = ("Context", "default_context", "B200Model", "load_model", "GetKeyPtsByAE", "GetKeyPtsFromRing",
        "GetKeyPtsFromRawFileName", "GetPatchesList", "GetFeaturesFromPatches", "GetFeaturesAtKeyPts",
        "SolveRT", "RANSAC4RT", "SolveRelativePose")


def __getattr__(name):
    if name in _API:
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
