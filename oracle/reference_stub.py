"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference modules from /root/reference.

Used only by ``tests/golden/make_golden.py`` (fixture generation in the build container,
where /root/reference is mounted) and by the ``not gpu`` oracle-validation tests that are
skipped when the mount is absent.  Nothing under ``caelo_b200/`` may import this.

The reference (SRainGit/CAE-LO) needs mayavi, matplotlib, cupy and numpy<1.20 aliases;
none exist here, so they are stubbed exactly as SURVEY.md Appendix C.2 describes:
  * ``mayavi``, ``mayavi.mlab``, ``matplotlib``, ``matplotlib.pyplot`` → dummy modules;
  * ``cupy`` → numpy's namespace + ``asnumpy`` + a STABLE ``argsort`` (the canonical
    tie rule of SURVEY quirk 9: ascending by (score, flat index));
  * ``np.bool`` / ``np.int`` aliases (Match.py:179,193).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_DIR = os.path.join(_ROOT, "baseline", "_ref")       # git-ignored working copy that travels to the GPU box
STAGED_FILES = ("Dirs.py", "Voxel.py", "SphericalRing.py", "Transformations.py", "Match.py", "MyICP.py",
                "PoseEstimation.py", "RefinePoses.py", "TrainedModels/SphericalRingPCRespondLayer.h5",
                "TrainedModels/EncoderModel4VoxelPatch.h5")


def _find_reference_dir() -> str:
    env = os.environ.get("CAELO_REFERENCE_DIR")
    if env:
        return env
    if os.path.isfile("/root/reference/Match.py"):
        return "/root/reference"
    return STAGED_DIR


REFERENCE_DIR = _find_reference_dir()


def stage(src: str = "/root/reference") -> bool:
    """Copy the UNMODIFIED reference modules of the hot path and the two inference .h5 files into
    ``baseline/_ref/`` (git-ignored, not gpurun-ignored: it travels to the GPU box, where /root/reference does not
    exist) — SURVEY.md §7 (i).  Called by ``__graft_entry__.build()`` when the reference tree is mounted."""
    import shutil
    if not os.path.isfile(os.path.join(src, "Match.py")):
        return False
    for rel in STAGED_FILES:
        dst = os.path.join(STAGED_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or os.path.getsize(dst) != os.path.getsize(os.path.join(src, rel)):
            shutil.copyfile(os.path.join(src, rel), dst)
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "Match.py"))


class _Dummy(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return None


_loaded = {}


def prepare():
    """Register the dependency stubs only (no reference module is imported)."""
    if not hasattr(np, "bool"):
        np.bool = bool  # type: ignore[attr-defined]
    if not hasattr(np, "int"):
        np.int = int  # type: ignore[attr-defined]
    for name in ("mayavi", "mayavi.mlab", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, _Dummy(name))
    cp = types.ModuleType("cupy")
    cp.__dict__.update({k: v for k, v in np.__dict__.items() if not k.startswith("__")})
    cp.asnumpy = np.asarray
    cp.bool = bool
    cp.argsort = lambda a, *args, **kw: np.argsort(a, kind="stable")
    sys.modules["cupy"] = cp


def load():
    """Return a dict of the imported reference modules (cached)."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_DIR)
    prepare()
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import Voxel  # noqa: F401  (≈6 s: builds 560k nested lists, Voxel.py:57-86)
        import SphericalRing  # noqa: F401
        import Transformations  # noqa: F401
        import Match  # noqa: F401
    finally:
        sys.path.remove(REFERENCE_DIR)
    _loaded.update(Voxel=Voxel, SphericalRing=SphericalRing, Transformations=Transformations,
                   Match=Match)
    return _loaded


# ---- the reference's hot path driven through its own functions (bench.py --impl reference, tests) -------------
class KerasStandIn:
    """``keras.models.load_model(path)`` stand-in for the CPU arm: Keras / TensorFlow are not installable here, so
    ``predict`` is the torch-CPU fp32 restatement of the two graphs with Keras' default
    ``batch_size=32``; the weights are read from the reference's own .h5 files when they are there."""

    def __init__(self, kind: str):
        from oracle import oracle
        self.kind = kind
        self.w = oracle.load_weights(kind)
        h5 = os.path.join(REFERENCE_DIR, "TrainedModels", "SphericalRingPCRespondLayer.h5" if kind == "respond"
                          else "EncoderModel4VoxelPatch.h5")
        if os.path.isfile(h5):
            sys.path.insert(0, _ROOT)
            from caelo_b200.h5weights import read_keras_weights     # a pure-python HDF5 reader, no device code
            self.w, _sha = read_keras_weights(h5)

    def predict(self, x, batch_size=32, verbose=0):
        import torch
        import torch.nn.functional as F
        from oracle import oracle
        if self.kind == "encoder":
            return oracle.encoder_predict(np.asarray(x, np.float32), self.w, batch_size)
        w = {k: torch.from_numpy(np.ascontiguousarray(v, np.float32)) for k, v in self.w.items()}
        k1 = w["conv2d_1/kernel:0"].permute(3, 2, 0, 1).contiguous()
        k2 = w["conv2d_2/kernel:0"].permute(3, 2, 0, 1).contiguous()
        xs = torch.from_numpy(np.ascontiguousarray(x, np.float32))
        outs = []
        with torch.no_grad():
            for i in range(0, xs.shape[0], batch_size):
                t = xs[i:i + batch_size].permute(0, 3, 1, 2)
                t = torch.relu(F.conv2d(t, k1, w["conv2d_1/bias:0"], padding=1))
                t = torch.relu(F.conv2d(t, k2, w["conv2d_2/bias:0"]))
                outs.append(t.permute(0, 2, 3, 1).contiguous().numpy())
        return np.concatenate(outs, 0)


_models = {}


def model(kind: str) -> KerasStandIn:
    if kind not in _models:
        _models[kind] = KerasStandIn(kind)
    return _models[kind]


def frame_stage(ring3, counter, vox0, vox1, vox2):
    """One frame through the reference's own code: RespondLayer.predict (SphericalRing.py:405-408) ->
    GetKeyPtsByAE (:113) -> GetPatchesList (Voxel.py:177) -> GetFeaturesFromPatches (Match.py:130).
    -> (KeyPts (n,3) f32, Features (n,60) f32)."""
    m = load()
    resp = model("respond").predict(np.asarray(ring3, np.float32)[None])[0]
    KeyPts, _KeyPixels, _Planar = m["SphericalRing"].GetKeyPtsByAE(ring3, counter, resp)
    KeyPts, PatchesList = m["Voxel"].GetPatchesList(KeyPts, vox0, vox1, vox2)
    Features = m["Match"].GetFeaturesFromPatches(model("encoder"), PatchesList)
    return np.asarray(KeyPts, np.float32), np.asarray(Features, np.float32)


def pair_stage(pair_id, KeyPts0, Features0, KeyPts1, Features1):
    """SolveRelativePose (Match.py:241) after ``np.random.seed(pair_id)`` (the harness convention of SURVEY §8d;
    the reference draws from the unseeded global stream) -> pose row [16]: R(9) T(3) isSuccess nInliers thr 0."""
    m = load()
    import contextlib
    import io as _io
    w = np.ones((KeyPts0.shape[0], 1), np.float32)
    np.random.seed(int(pair_id))
    with contextlib.redirect_stdout(_io.StringIO()):                  # the reference prints its progress
        R, T, ok, i0, _i1, thr = m["Match"].SolveRelativePose(KeyPts0, Features0, w, KeyPts1, Features1, w.copy())
    return np.r_[np.asarray(R, np.float32).ravel(), np.asarray(T, np.float32).ravel(), float(ok), len(i0), thr, 0.0]
