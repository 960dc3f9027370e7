"""The drop-in boundary: shadow modules named like the reference's (`Match`, `Voxel`,
`SphericalRing`) + a keras shim, so that the reference's drivers pick up the B200 hot path through their own
`from X import *` lines.  CPU only: no compute call is made."""
import subprocess
import sys
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code):
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_standalone_shadow_modules_and_keras_shim():
    out = _run("""
from caelo_b200 import dropin, api, odometry
mods = dropin.install()
from Match import *
from Voxel import *
from SphericalRing import *
assert SolveRelativePose is api.SolveRelativePose and RANSAC4RT is api.RANSAC4RT and SolveRT is api.SolveRT
assert GetFeaturesFromPatches is api.GetFeaturesFromPatches and GetPatchesList is api.GetPatchesList
assert Voxelization is api.Voxelization and ProjectPC2SphericalRing is api.ProjectPC2SphericalRing
assert GetKeyPtsByAE is api.GetKeyPtsByAE and GetKeyPtsFromRawFileName is api.GetKeyPtsFromRawFileName
assert ExtendKeyPtsInShpericalRing is api.ExtendKeyPtsInShpericalRing
assert LoadVoxelModelAndKeyPts is odometry.LoadVoxelModelAndKeyPts
from MyICP import *
assert ICP is api.ICP and GetPtsInliners is api.GetPtsInliners and ICP_Pt2PtAndPt2Plane is api.ICP_Pt2PtAndPt2Plane
assert (nLines, ImgH, ImgW, CropWidth_SphericalRing, Channels4AE) == (64, 69, 1800, 8, [0, 1, 2])
assert abs(VisibleLength - 99.84) < 1e-12 and VoxelSizes == [0.02, 0.16, 0.64] and PatchSize == 16
assert strVoxelPatchEncoderPath.endswith('EncoderModel4VoxelPatch.h5')
import keras
from keras.models import load_model, Model
assert load_model is api.load_model and Model is api.B200Model
print('ok')
""")
    assert out.strip().endswith("ok")


@pytest.mark.reference
def test_shadowing_the_unmodified_reference_modules():
    """With the reference tree present its own modules are executed and ONLY the hot functions are replaced —
    in every namespace that star-imported them (Match.py does `from Voxel import *`)."""
    from oracle import reference_stub
    if not reference_stub.available():
        pytest.skip("reference tree not mounted")
    out = _run("""
from oracle import reference_stub
reference_stub.prepare()                       # mayavi / cupy / np.bool stand-ins the reference needs to import
from caelo_b200 import dropin, api
mods = dropin.install(reference_stub.REFERENCE_DIR)
import Match, Voxel, SphericalRing
assert Match.__file__.startswith(reference_stub.REFERENCE_DIR)
assert Match.SolveRelativePose is api.SolveRelativePose and Match.GetPatchesList is api.GetPatchesList
assert SphericalRing.GetKeyPtsByAE is api.GetKeyPtsByAE and SphericalRing.Voxelization is api.Voxelization
assert Voxel.GetPatchesList is api.GetPatchesList
import MyICP
assert MyICP.ICP is api.ICP and MyICP.SolveRT is api.SolveRT and MyICP.ICP_Pt2PtAndPt2Plane is api.ICP_Pt2PtAndPt2Plane
# untouched helpers of the reference are still the reference's
assert Match.GetKeyVoxelsAroundKeyPts.__module__ == 'Match' and callable(SphericalRing.LocateKeyPixels)
assert Voxel.nBlocksL == 156 and SphericalRing.ImgW == 1800
print('ok')
""")
    assert out.strip().endswith("ok")
