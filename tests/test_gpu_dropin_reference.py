"""The drop-in boundary exercised on the GPU under the reference's OWN drivers: the
unmodified reference modules from ``baseline/_ref`` (staged by ``__graft_entry__.build()``; /root/reference in the
build container) are executed, ``dropin.install`` replaces only the hot functions + Keras, and the statements of the
``Match.py`` two-frame demo (:313-349, BASELINE configs[0]) and ``PoseEstimation.GetRelativePoseBetween2Frames``
(:152-169) run unmodified against a DemoData-shaped tree materialised from the committed fixtures.  Results are
compared with the oracle and with what the reference's own functions compute on the CPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_data as G

pytestmark = [pytest.mark.gpu, pytest.mark.reference]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tree(tmp, tags):
    """<tmp>/velodyne/sequences/SS/{velodyne,SphericalRing,VoxelModel,KeyPts}/NNNNNN.bin[.mat] as the reference's
    loaders expect them (Match.py:46-61, SphericalRing.py:394-401)."""
    from scipy import io
    files = []
    for tag in tags:
        seq, name = tag.split("_")
        f = G.frame(tag)
        base = os.path.join(tmp, "velodyne", "sequences", seq)
        for sub in ("velodyne", "SphericalRing", "VoxelModel", "KeyPts"):
            os.makedirs(os.path.join(base, sub), exist_ok=True)
        raw = os.path.join(base, "velodyne", name + ".bin")
        open(raw, "wb").close()                                      # the demo only uses the NAME of the raw file here
        io.savemat(os.path.join(base, "SphericalRing", name + ".bin.mat"),
                   {"SphericalRing": f["ring5"], "GridCounter": f["counter"]})
        io.savemat(os.path.join(base, "VoxelModel", name + ".bin.mat"),
                   {"AllVoxels0": f["vox0"], "AllVoxels1": f["vox1"], "AllVoxels2": f["vox2"]})
        io.savemat(os.path.join(base, "KeyPts", name + ".bin.mat"), {"KeyPts": f["golden_KeyPts"]})
        files.append(raw)
    return files


DEMO = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, ROOT)
from oracle import reference_stub
reference_stub.prepare()                                  # mayavi / matplotlib / cupy stand-ins the reference imports
from caelo_b200 import dropin, api
mods = dropin.install(reference_stub.REFERENCE_DIR)
os.chdir(reference_stub.REFERENCE_DIR)                    # the model paths of Dirs.py:29-30 are relative
sys.path.insert(0, reference_stub.REFERENCE_DIR)
from Match import *                                       # what the demo's module body sees
assert SolveRelativePose is api.SolveRelativePose and GetPatchesList is api.GetPatchesList
FileName0, FileName1 = FILES
# ---- Match.py:311-349, statement for statement (timers and prints dropped) ----
import keras
from keras.models import load_model
PatchEncoder = load_model(strVoxelPatchEncoderPath)
KeyPts0, AllVoxels00, AllVoxels01, AllVoxels02 = LoadVoxelModelAndKeyPts(FileName0)
KeyPts1, AllVoxels10, AllVoxels11, AllVoxels12 = LoadVoxelModelAndKeyPts(FileName1)
RespondLayer = load_model(strRespondNetModelPath)
KeyPts0, KeyPixels0, PlanarPts0 = GetKeyPtsFromRawFileName(FileName0, RespondLayer)
KeyPts1, KeyPixels1, PlanarPts1 = GetKeyPtsFromRawFileName(FileName1, RespondLayer)
KeyPts0, PatchesList0 = GetPatchesList(KeyPts0, AllVoxels00, AllVoxels01, AllVoxels02)
KeyPts1, PatchesList1 = GetPatchesList(KeyPts1, AllVoxels10, AllVoxels11, AllVoxels12)
Features0 = GetFeaturesFromPatches(PatchEncoder, PatchesList0)
Features1 = GetFeaturesFromPatches(PatchEncoder, PatchesList1)
Weights0 = np.ones((KeyPts0.shape[0],1),dtype=np.float32)
Weights1 = np.ones((KeyPts1.shape[0],1),dtype=np.float32)
np.random.seed(0)                                         # harness seed (the reference draws from the unseeded stream)
R, T, score, inliersIdx0, inliersIdx1, residualThreshold = SolveRelativePose(KeyPts0, Features0, Weights0, KeyPts1, Features1, Weights1)
pairs0 = KeyPts0[inliersIdx0,:]
pairs1 = KeyPts1[inliersIdx1,:]
# ---- PoseEstimation.py:152-169 through the reference's own driver module ----
import PoseEstimation
assert PoseEstimation.SolveRelativePose is api.SolveRelativePose
PoseEstimation.listKeyPtsData = [[KeyPts0, Features0, Weights0], [KeyPts1, Features1, Weights1]]
PoseEstimation.inliersData = []
np.random.seed(0)
relativeR, relativeT, isSuccess, nInliers, thr = PoseEstimation.GetRelativePoseBetween2Frames(0, 1)
np.savez(OUT, KeyPts0=KeyPts0, KeyPts1=KeyPts1, KeyPixels0=KeyPixels0, Features0=Features0, Features1=Features1, R=R, T=T,
         score=score, idx0=inliersIdx0, idx1=inliersIdx1, thr=residualThreshold, relativeR=relativeR, relativeT=relativeT,
         isSuccess=isSuccess, nInliers=nInliers, inliersData_idx0=PoseEstimation.inliersData[0][2],
         launches=api.default_context().launches, patch_sum=float(PatchesList0[1].sum()))
print('demo ok')
'''


@pytest.mark.parametrize("seq", ["00", "01"])
def test_match_demo_and_pose_driver_through_the_dropin(oracle_mod, tmp_path, seq):
    import torch
    from oracle import reference_stub
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not reference_stub.available():
        pytest.skip("reference tree neither mounted nor staged under baseline/_ref")
    tags = G.PAIRS[seq]
    files = _tree(str(tmp_path), tags)
    out = str(tmp_path / "demo.npz")
    code = "ROOT = %r\nFILES = %r\nOUT = %r\n" % (ROOT, files, out) + DEMO
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "demo ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    z = np.load(out)
    assert int(z["launches"]) > 20                                       # the hot path ran on the device
    # (1) key points: the reference's own GetKeyPtsByAE on the 5-channel ring (reference-run fixture), bit for bit
    f0, f1 = G.frame(tags[0]), G.frame(tags[1])
    assert np.array_equal(z["KeyPts0"], G.refrun(tags[0])["keypts_ring5_i32"])
    assert np.array_equal(z["KeyPixels0"], G.refrun(tags[0])["keypix_ring5_i32"])
    assert np.array_equal(z["KeyPts1"], G.refrun(tags[1])["keypts_ring5_i32"])
    # (2) descriptors: oracle (torch-CPU restatement of the Keras graph on the oracle's patches) within the contract
    for kp, ft, f in ((z["KeyPts0"], z["Features0"], f0), (z["KeyPts1"], z["Features1"], f1)):
        _, pl = oracle_mod.get_patches_list(kp, f["vox0"], f["vox1"], f["vox2"])
        want = oracle_mod.get_features_from_patches(pl)
        assert (np.abs(ft - want) <= 1e-4 * np.abs(want) + 1e-5).all()
    # (3) pose: the oracle's SolveRelativePose on the same key points + descriptors with the same seed, bit for bit,
    #     and the reference's own SolveRelativePose (unmodified Match.py on the CPU) within 1e-4
    np.random.seed(0)
    Ro, To, oko, i0, i1, thro = oracle_mod.solve_relative_pose(z["KeyPts0"], z["Features0"], None, z["KeyPts1"], z["Features1"], None)
    assert np.array_equal(z["R"], Ro) and np.array_equal(z["T"], To) and bool(z["score"]) == oko and float(z["thr"]) == thro
    assert np.array_equal(z["idx0"], i0) and np.array_equal(z["idx1"], i1)
    row = reference_stub.pair_stage(0, z["KeyPts0"], z["Features0"], z["KeyPts1"], z["Features1"])
    assert int(row[13]) == len(i0) and bool(row[12]) == oko
    assert np.abs(row[:9] - z["R"].ravel()).max() <= 1e-4 and np.abs(row[9:12] - z["T"].ravel()).max() <= 1e-4 * max(1.0, np.abs(row[9:12]).max())
    # (4) the PoseEstimation.py driver function returned the same thing and recorded the inliers (:165)
    assert np.array_equal(z["relativeR"], z["R"]) and np.array_equal(z["relativeT"], z["T"].reshape(3, 1))
    assert bool(z["isSuccess"]) == oko and int(z["nInliers"]) == len(i0) and np.array_equal(z["inliersData_idx0"], i0)
