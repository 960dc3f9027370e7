"""Drop-in shadow modules for the reference's own drivers.

The reference's scripts do ``from Match import *`` / ``from Voxel import *`` / ``from SphericalRing
import *`` (PoseEstimation.py:20-22, BatchPreprocess.py:19-22) and import Keras lazily inside functions
(Match.py:311-313, PoseEstimation.py:71-73, BatchPreprocess.py:166-168).  ``install()`` registers modules
of those names in ``sys.modules``:

  * with ``reference_dir`` the ORIGINAL module is executed first under its own name (so every helper,
    constant and path the drivers use is there) and the hot-path functions are then replaced by the
    B200 ones — the reference keeps working, only the hot path changes;
  * without it (the GPU box has no reference tree) the modules carry the module-level constants the
    drivers read plus the hot-path functions.

``keras`` / ``keras.models`` are registered as shims whose ``load_model(path)`` returns a ``B200Model``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from typing import Optional

from . import api, odometry

# names each shadow module overrides (reference symbol -> replacement)
HOT = {
    "Voxel": {"Voxelization": api.Voxelization, "GetPatchesList": api.GetPatchesList},
    "SphericalRing": {"ProjectPC2SphericalRing": api.ProjectPC2SphericalRing, "GetKeyPtsByAE": api.GetKeyPtsByAE,
                      "GetKeyPtsFromRawFileName": api.GetKeyPtsFromRawFileName,
                      "ExtendKeyPtsInShpericalRing": api.ExtendKeyPtsInShpericalRing},
    "Match": {"GetFeaturesFromPatches": api.GetFeaturesFromPatches, "SolveRT": api.SolveRT, "RANSAC4RT": api.RANSAC4RT,
              "SolveRelativePose": api.SolveRelativePose, "LoadVoxelModelAndKeyPts": odometry.LoadVoxelModelAndKeyPts,
              "LoadKeyPtsAndFeatures": odometry.LoadKeyPtsAndFeatures},
    "MyICP": {"ICP": api.ICP, "GetPtsInliners": api.GetPtsInliners, "GetPlanarPtsInliners": api.GetPlanarPtsInliners,
              "ICP_Pt2PtAndPt2Plane": api.ICP_Pt2PtAndPt2Plane},      # f4 (RefinePoses.py does `from MyICP import *`)
}

# module-level constants the drivers read after ``from X import *`` (Voxel.py:15-52, SphericalRing.py:28-62, Dirs.py:29-30)
CONSTANTS = {
    "Voxel": dict(VoxelSize=api.VoxelSize, PatchSize=api.PatchSize, Scales=api.Scales, VoxelSizes=api.VoxelSizes,
                  VisibleLength=api.VisibleLength, VisibleWidth=api.VisibleWidth, VisibleHeight=api.VisibleHeight,
                  PatchRadius=8, BlockRealSize=1.28, BlockSize=64),
    "SphericalRing": dict(nLines=api.nLines, ImgH=api.ImgH, ImgW=api.ImgW, NumChannels=5,
                          CropWidth_SphericalRing=api.CropWidth_SphericalRing, Channels4AE=api.Channels4AE,
                          SafeEdgeWidth4Top=5, Size4FilterTopEdge=8),
    "Match": dict(nFixedKeyPts=api.nFixedKeyPts,
                  strRespondNetModelPath="./TrainedModels/SphericalRingPCRespondLayer.h5",
                  strVoxelPatchEncoderPath="./TrainedModels/EncoderModel4VoxelPatch.h5"),
    "MyICP": dict(),
}


def _load_original(name: str, reference_dir: str) -> types.ModuleType:
    path = os.path.join(reference_dir, name + ".py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod          # the reference modules import each other by these names
    spec.loader.exec_module(mod)
    return mod


def install_keras_shim():
    keras = types.ModuleType("keras")
    models = types.ModuleType("keras.models")
    models.load_model = api.load_model
    models.Model = api.B200Model
    keras.models = models
    sys.modules["keras"] = keras
    sys.modules["keras.models"] = models
    return keras


def install(reference_dir: Optional[str] = None, keras_shim: bool = True):
    """Register the shadow modules (and the Keras shim).  Returns {name: module}."""
    out = {}
    if reference_dir:
        sys.path.insert(0, reference_dir)
    try:
        for name in ("Voxel", "SphericalRing", "Match", "MyICP"):       # dependency order of the reference
            if reference_dir and os.path.isfile(os.path.join(reference_dir, name + ".py")):
                mod = _load_original(name, reference_dir)
            else:
                mod = types.ModuleType(name)
                mod.__dict__.update(CONSTANTS[name])
                sys.modules[name] = mod
            for sym, fn in HOT[name].items():
                setattr(mod, sym, fn)
            out[name] = mod
        # ``from Voxel import *`` inside SphericalRing / Match copied the ORIGINAL functions into their
        # namespaces before the override: patch those copies too
        for name, mod in out.items():
            for other in HOT.values():
                for sym, fn in other.items():
                    if hasattr(mod, sym):
                        setattr(mod, sym, fn)
    finally:
        if reference_dir:
            sys.path.remove(reference_dir)
    if keras_shim:
        out["keras"] = install_keras_shim()
    return out
