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

REFERENCE_DIR = os.environ.get("CAELO_REFERENCE_DIR", "/root/reference")


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
