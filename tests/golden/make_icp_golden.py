"""Fixture for the f4 row (ICP on the extended key points): run the UNMODIFIED reference ``MyICP.ICP``
(MyICP.py:28-73: sklearn 1-NN + numpy float32 SolveRT) in the build container on the DemoData pairs and store
its result.  Inputs are NOT stored: the tests rebuild them from the committed frame fixtures with the oracle
(respond -> GetKeyPtsByAE -> ExtendKeyPtsInShpericalRing, all pinned elsewhere), exactly as this script does.

    python tests/golden/make_icp_golden.py        ->  tests/golden/icp_SS.npz
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_data as G  # noqa: E402
from oracle import oracle, reference_stub  # noqa: E402


def extended_keypoints(tag):
    """ExtendedKeyPts of a demo frame as BatchPreprocess.py:136-141 produces them (5-channel ring, int32 counter)."""
    f = G.frame(tag)
    resp = oracle.respond_predict(f["ring3"][None])[0]
    _kp, px = oracle.select_keypoints(f["ring3"], f["counter_i8"], resp)
    return oracle.extend_keypoints(f["ring5"], f["counter"].copy(), px)


def icp_inputs(seq):
    """(KeyPts0, KeyPts1_): the extended key points of the pair, frame 1 moved by the odometry pose first
    (RefinePoses.py:287: KeyPts1_ = float32(R KeyPts1^T + T)^T) — here the seed-0 SolveRelativePose fixture."""
    t0, t1 = G.PAIRS[seq]
    k0, k1 = extended_keypoints(t0), extended_keypoints(t1)
    p = G.pose(seq)
    R, T = p["R_0"], p["T_0"]
    k1_ = np.array((np.dot(R, k1.T) + T.reshape(3, 1)).T, dtype=np.float32)
    return k0, k1, k1_


def planar_inputs(k, n_max=3000):
    """A deterministic stand-in for the planar points the shipped pipeline never produces (SphericalRing.py:219):
    every 4th extended key point with a plausible unit normal — up for ground points, towards the sensor otherwise."""
    p = k[::4][:n_max].astype(np.float32)
    n = np.zeros_like(p)
    ground = p[:, 2] < -1.3
    n[ground, 2] = 1.0
    h = -p[~ground, 0:2]
    n[~ground, 0:2] = h / np.linalg.norm(h, axis=1, keepdims=True)
    return np.ascontiguousarray(np.c_[p, n].astype(np.float32))


def main():
    reference_stub.load()
    sys.path.insert(0, reference_stub.REFERENCE_DIR)
    import MyICP  # noqa: E402  (unmodified reference)
    for seq in G.PAIRS:
        k0, k1, k1_ = icp_inputs(seq)
        out = {}
        for name, pc1, kw in (("aligned", k1_, {}), ("raw", k1, {}),
                              ("tight", k1_, dict(inlierThreshold=0.3, smallShiftThreshold=0.1, ep=0.01))):
            with contextlib.redirect_stdout(io.StringIO()) as buf:
                R, T, ok = MyICP.ICP(k0.copy(), pc1.copy(), **kw)
            out["R_" + name], out["T_" + name], out["ok_" + name] = R, T, ok
            out["log_" + name] = buf.getvalue().strip()
            print(seq, name, k0.shape, pc1.shape, ok, buf.getvalue().strip())
        # ICP_Pt2PtAndPt2Plane (MyICP.py:127-201) with the arguments of RefinementCore (RefinePoses.py:293-296);
        # frame 1's planar coordinates are moved by the odometry pose like its key points (:289-290)
        pl0, pl1 = planar_inputs(k0), planar_inputs(k1)
        p = G.pose(seq)
        pl1[:, 0:3] = np.array((np.dot(p["R_0"], pl1[:, 0:3].T) + p["T_0"].reshape(3, 1)).T, dtype=np.float32)
        np.random.seed(7)
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            R, T, ok = MyICP.ICP_Pt2PtAndPt2Plane(k0.copy(), k1_.copy(), pl0.copy(), pl1.copy(), maxIterTimes=50,
                                                  minIterTimes=20 - 1, inlierThreshold0=0.5, decay_rate0=0.9,
                                                  inlierThreshold1=5.0, decay_rate1=0.9, smallShiftThreshold=0.1, ep=0.001)
        out["R_plane"], out["T_plane"], out["ok_plane"], out["log_plane"] = R, T, ok, buf.getvalue().strip()
        out["next_random_plane"] = np.random.random()
        print(seq, "plane", pl0.shape, pl1.shape, ok, buf.getvalue().strip())
        np.savez_compressed(os.path.join(HERE, "icp_%s.npz" % seq), **out)


if __name__ == "__main__":
    main()
