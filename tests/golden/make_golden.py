"""Generate the committed fixtures under tests/golden/ (run ONCE in the build container,
where /root/reference is mounted; the GPU box has neither the reference nor its DemoData).

    python tests/golden/make_golden.py

What is produced and from what:
  caelo_b200/weights/{respond,encoder}.npz   the float32 datasets of the two shipped Keras .h5
                                             files (TrainedModels/), read with caelo_b200.h5weights
  frame_SS_NNNNNN.npz   inputs of the hot path for the four DemoData scans, produced by the
                        UNMODIFIED reference functions ProjectPC2SphericalRing (SphericalRing.py:72)
                        and Voxelization (Voxel.py:100, float64 guard — SURVEY quirk 7), stored
                        sparsely, plus the reference's own golden KeyPts/Features (.mat in the zip)
  refrun_SS_NNNNNN.npz  outputs of the UNMODIFIED reference GetKeyPtsByAE (both ring variants)
                        fed with the oracle's response image, and of GetPatchesList at the golden
                        keypoints (bit-packed)
  pose_SS.npz           outputs of the UNMODIFIED reference SolveRelativePose on the golden
                        descriptors of each demo pair for np.random.seed(0..2)
  usip_*.npz            DemoData third-party 128-D descriptors (nn-match KAT inputs) + scipy result
  scan_SS_NNNNNN.npz    (``--scans``) two raw DemoData scans (N,4) f32 with the block list / block
                        offsets / in-block voxels the UNMODIFIED reference Voxelization returns for
                        them; together with frame_*.npz (ring, counter, AllVoxels0/1/2 from the same
                        reference run) they pin the f1/f2 pre-stages
"""
from __future__ import annotations

import io
import json
import os
import sys
import zipfile

import numpy as np
import scipy.io
from scipy.spatial.distance import cdist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from caelo_b200.h5weights import read_keras_weights, read_model_config, layer_summary  # noqa: E402
from oracle import oracle, reference_stub  # noqa: E402

REF = reference_stub.REFERENCE_DIR
ZIP = os.path.join(REF, "DemoData", "KITTI_odometry.zip")
FRAMES = [("00", 0), ("00", 1), ("01", 495), ("01", 496)]


def export_weights():
    out = os.path.join(ROOT, "caelo_b200", "weights")
    os.makedirs(out, exist_ok=True)
    meta = {}
    for name, fn in (("respond", "SphericalRingPCRespondLayer.h5"), ("encoder", "EncoderModel4VoxelPatch.h5")):
        path = os.path.join(REF, "TrainedModels", fn)
        w, sha = read_keras_weights(path)
        np.savez(os.path.join(out, name + ".npz"), **w)
        meta[name] = {"source": "TrainedModels/" + fn, "sha256": sha,
                      "layers": [list(map(str, l[:3])) for l in layer_summary(read_model_config(path))]}
    with open(os.path.join(out, "weights.json"), "w") as f:
        json.dump(meta, f, indent=1)


SCAN_FRAMES = [("00", 0), ("01", 495)]


def export_scans():
    ref = reference_stub.load()
    VX = ref["Voxel"]
    z = zipfile.ZipFile(ZIP)
    for seq, fr in SCAN_FRAMES:
        pc = np.frombuffer(z.read("KITTI_odometry/velodyne/sequences/%s/velodyne/%06d.bin" % (seq, fr)),
                           np.float32).reshape(-1, 4)
        vox = VX.Voxelization(pc.astype(np.float64))
        np.savez_compressed(os.path.join(HERE, "scan_%s_%06d.npz" % (seq, fr)), pc=pc,
                            avlBlocksList=vox[3], cntVoxelsLength=vox[4], AllVoxels=vox[5])
        print("scan", seq, fr, pc.shape, vox[3].shape, vox[5].shape)


def main():
    if "--scans" in sys.argv:
        return export_scans()
    export_weights()
    ref = reference_stub.load()
    SR, VX, MT = ref["SphericalRing"], ref["Voxel"], ref["Match"]
    z = zipfile.ZipFile(ZIP)

    def zread(name):
        return z.read("KITTI_odometry/" + name)

    golden = {}
    for seq, fr in FRAMES:
        tag = "%s_%06d" % (seq, fr)
        pc = np.frombuffer(zread("velodyne/sequences/%s/velodyne/%06d.bin" % (seq, fr)), np.float32).reshape(-1, 4)
        ring, counter = SR.ProjectPC2SphericalRing(pc)
        vox = VX.Voxelization(pc.astype(np.float64))
        av0, av1, av2 = vox[6], vox[7], vox[8]
        mat = scipy.io.loadmat(io.BytesIO(zread("velodyne/sequences/%s/Features/%06d.bin.mat" % (seq, fr))))
        occ = np.flatnonzero(counter.reshape(-1) > 0).astype(np.int32)
        np.savez_compressed(
            os.path.join(HERE, "frame_%s.npz" % tag),
            ring_idx=occ, ring_val=ring.reshape(-1, 5)[occ], counter_val=counter.reshape(-1)[occ].astype(np.int32),
            vox0=av0, vox1=av1, vox2=av2,
            golden_KeyPts=mat["KeyPts"], golden_Features=mat["Features"], n_points=np.int64(pc.shape[0]))
        golden[tag] = (mat["KeyPts"], mat["Features"])

        # reference-run outputs
        x = ring[0:64, 0:1792, 0:3][None]
        resp = oracle.respond_predict(x)[0]
        kp5, px5, _ = SR.GetKeyPtsByAE(ring, counter, resp)
        ring3 = np.array(ring[0:64, 0:1792, [0, 1, 2]], dtype=np.float32)
        kp3, px3, _ = SR.GetKeyPtsByAE(ring3, counter.astype(np.int8), resp)
        _, plist = VX.GetPatchesList(mat["KeyPts"], av0, av1, av2)
        packed = np.stack([np.packbits(p.reshape(p.shape[0], -1) > 0, axis=1) for p in plist])
        np.savez_compressed(os.path.join(HERE, "refrun_%s.npz" % tag),
                            keypix_ring5_i32=px5.astype(np.int64), keypts_ring5_i32=kp5,
                            keypix_ring3_i8=px3.astype(np.int64), keypts_ring3_i8=kp3,
                            patches_packed=packed)
        print(tag, "pts", pc.shape[0], "occ", occ.size, "vox", av0.shape[0], av1.shape[0], av2.shape[0],
              "kp5", kp5.shape, "kp3", kp3.shape)

    export_scans()

    for seq, f0, f1 in (("00", 0, 1), ("01", 495, 496)):
        k0, c0 = golden["%s_%06d" % (seq, f0)]
        k1, c1 = golden["%s_%06d" % (seq, f1)]
        out = {}
        D = cdist(c0, c1, metric="euclidean")
        out["pair_idx"] = np.argmin(D, axis=0).astype(np.int64)
        for seed in range(3):
            np.random.seed(seed)
            R, T, ok, i0, i1, thr = MT.SolveRelativePose(k0, c0, None, k1, c1, None)
            out["R_%d" % seed] = np.asarray(R)
            out["T_%d" % seed] = np.asarray(T)
            out["ok_%d" % seed] = np.bool_(ok)
            out["idx0_%d" % seed] = np.asarray(i0, np.int64)
            out["idx1_%d" % seed] = np.asarray(i1, np.int64)
            out["thr_%d" % seed] = np.float64(thr)
            out["next_random_%d" % seed] = np.float64(np.random.random())  # stream position probe
        np.savez_compressed(os.path.join(HERE, "pose_%s.npz" % seq), **out)

    # third-party 128-D descriptors: nn-match KAT
    for seq, f0, f1 in (("00", 0, 1), ("01", 495, 496)):
        d0 = np.frombuffer(zread("output_USIP/Descriptors/%s/%06d.bin" % (seq, f0)), np.float32).reshape(-1, 128)
        d1 = np.frombuffer(zread("output_USIP/Descriptors/%s/%06d.bin" % (seq, f1)), np.float32).reshape(-1, 128)
        idx = np.argmin(cdist(d0, d1, metric="euclidean"), axis=0).astype(np.int64)
        np.savez_compressed(os.path.join(HERE, "usip_%s.npz" % seq), d0=d0, d1=d1, pair_idx=idx)


if __name__ == "__main__":
    main()
