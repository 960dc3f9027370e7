"""GPU tests of the odometry file plumbing (caelo_b200/odometry.py) and of the BASELINE
configs that are parity cases rather than bench lines: a whole (short) sequence through
``estimate_sequence`` — files in the reference's formats, sharded like configs[2] — and the descriptor
NN-match microbench shape of configs[3] (up to 16k x 16k x 128)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from caelo_b200 import api as a
    a.default_context()
    return a


@pytest.fixture(scope="module")
def sequence(tmp_path_factory):
    """Seven synthetic KITTI-shaped scans written as velodyne/NNNNNN.bin + a calib_.txt with a non-trivial Tr."""
    from caelo_b200 import synth
    d = synth.make_frames(7, seed=21)
    root = tmp_path_factory.mktemp("seq")
    raw = root / "velodyne"
    raw.mkdir()
    for i, pc in enumerate(d["scans"]):
        np.asarray(pc, np.float32).tofile(str(raw / ("%06d.bin" % i)))
    calib = np.zeros((5, 12))
    calib[:4] = np.eye(3, 4).ravel()
    a = 0.02
    Tr = np.array([[np.cos(a), -np.sin(a), 0, 0.1], [np.sin(a), np.cos(a), 0, -0.2], [0, 0, 1, 0.3]])
    calib[4] = Tr.ravel()
    np.savetxt(str(root / "calib_.txt"), calib)
    return dict(root=str(root), raw=str(raw), scans=d["scans"], Tr=np.asarray(Tr, np.float32))


def test_estimate_sequence_matches_per_pair_api_and_writes_reference_files(api, sequence, tmp_path):
    """PoseEstimation.py:185-310 for one sequence: rel poses == the per-pair functions (ProjectPC2SphericalRing ->
    GetKeyPtsByAE -> Voxelization -> GetPatchesList -> encoder -> SolveRelativePose with np.random.seed(pair)),
    poses_/SS.txt == the reference's chaining recurrence, Features / InliersIdx .mat files as the reference writes."""
    from scipy import io
    from caelo_b200 import odometry, pipeline
    Tr = odometry.read_calib_Tr(os.path.join(sequence["root"], "calib_.txt"))
    assert np.allclose(Tr, sequence["Tr"])
    poses_path = str(tmp_path / "poses_" / "00.txt")
    fdir, idir = str(tmp_path / "Features"), str(tmp_path / "InliersIdx")
    poses, rel = odometry.estimate_sequence(sequence["raw"], Tr=Tr, poses_path=poses_path, features_dir=fdir,
                                            inliers_dir=idir, batch_pairs=4)
    F = len(sequence["scans"])
    assert poses.shape == (F, 12) and rel.shape == (F - 1, 16)
    assert np.allclose(np.loadtxt(poses_path), poses)
    assert np.array_equal(poses, pipeline.chain_poses(rel, Tr))
    # per-pair API on the same scans
    enc = api.load_model(os.path.join(api.WEIGHT_DIR, "encoder.npz"))
    kps, feats = [], []
    for pc in sequence["scans"]:
        ring, counter = api.ProjectPC2SphericalRing(pc)
        kp, _px, _ = api.GetKeyPtsFromRing(np.ascontiguousarray(ring[0:64, 0:1792, 0:3]), counter.astype(np.int8))
        vox = api.Voxelization(pc)
        _, pl = api.GetPatchesList(kp, *vox[6:9])
        kps.append(kp)
        feats.append(api.GetFeaturesFromPatches(enc, pl))
    for i in range(F - 1):
        np.random.seed(i)
        R, T, ok, i0, i1, thr = api.SolveRelativePose(kps[i], feats[i], None, kps[i + 1], feats[i + 1], None)
        assert bool(rel[i, 12]) == ok and int(rel[i, 13]) == len(i0) and abs(rel[i, 14] - thr) < 1e-6
        assert np.array_equal(rel[i, :9].reshape(3, 3), np.asarray(R, np.float32))
        assert np.array_equal(rel[i, 9:12], np.asarray(T, np.float32).ravel())
        m = io.loadmat(os.path.join(idir, "%06d-%06d.bin.mat" % (i, i + 1)))
        assert int(m["iFrame0"].item()) == i and int(m["iFrame1"].item()) == i + 1
        assert np.array_equal(m["inliersIdx0"].ravel(), i0) and np.array_equal(m["inliersIdx1"].ravel(), i1)
    for i in range(F):
        m = io.loadmat(os.path.join(fdir, "%06d.bin.mat" % i))
        assert np.array_equal(m["KeyPts"], kps[i]) and np.array_equal(m["Features"], feats[i])
        assert m["Weights"].shape == (1024, 1) and (m["Weights"] == 1).all()
    # the readers of Match.py:65-72 see what was written
    k, f, w = odometry.LoadKeyPtsAndFeatures(os.path.join(str(tmp_path), "velodyne", "000003.bin"))
    assert np.array_equal(k, kps[3]) and np.array_equal(f, feats[3])


def test_estimate_sequence_sharded_equals_single_rank(api, sequence):
    """configs[2] semantics without a process group: the pair ranges of two 'ranks' (one-frame halo recomputed)
    concatenate to exactly the single-rank result, whatever the batch size."""
    from caelo_b200 import odometry
    _, rel_all = odometry.estimate_sequence(sequence["raw"], scans=sequence["scans"], batch_pairs=32)
    parts = [odometry.estimate_sequence(sequence["raw"], scans=sequence["scans"], batch_pairs=2, rank=r, world=2)[1]
             for r in range(2)]
    assert np.array_equal(np.concatenate(parts, 0), rel_all)


def test_long_drive_sharded_equals_single_rank_and_stays_on_track(api):
    """configs[2] at a size that exercises many batches: a 258-frame synthetic drive (257 pairs) from scans resident in
    HBM — the pair ranges of 2 and of 8 'ranks' (every rank holding only its own frames + the one-frame halo) concatenate
    to exactly the single-rank result; every pair registers, and the chained trajectory stays close to the known motion."""
    import torch
    from caelo_b200 import odometry, pipeline, synth
    F = 258
    pts, off = synth.make_scans(F, seed=7, device="cuda")
    pipe = pipeline.OdometryPipeline(api.default_context())
    poses, rel = odometry.estimate_sequence(stacked=(pts, off), batch_pairs=32, pipe=pipe)
    assert rel.shape == (F - 1, 16) and (rel[:, 12] == 1).all()
    for world in (2, 8):
        parts = []
        for r in range(world):
            lo, hi = pipeline.shard_pairs(F - 1, r, world)
            sl = (pts[int(off[lo]):int(off[hi + 1])], off[lo:hi + 2] - off[lo])          # this rank's frames only
            parts.append(odometry.estimate_sequence(stacked=sl, stacked_first_frame=lo, n_frames=F, batch_pairs=32, rank=r,
                                                    world=world, pipe=pipe)[1])
        assert np.array_equal(np.concatenate(parts, 0), rel), world
    # host-resident input (pinned, streamed batch by batch) gives the same rows
    host = torch.empty(pts.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(pts)
    _, rel_h = odometry.estimate_sequence(stacked=(host, off), batch_pairs=32, pipe=pipe)
    assert np.array_equal(rel_h, rel)
    gt = _ground_truth(F)
    rre, rte = _pose_errors(poses.astype(np.float64), gt)
    assert rre.max() < 1.0 and rte.max() < 0.5                        # the reference's success criterion, every pair
    assert np.linalg.norm(poses[-1].reshape(3, 4)[:, 3] - gt[-1].reshape(3, 4)[:, 3]) < 0.1 * 0.7 * (F - 1)


def test_preprocess_sequence_writes_what_the_loaders_read(api, sequence, oracle_mod):
    """BatchPreprocess.py:44-67 / BatchVoxelization.py:42-64 outputs for two frames: .mat contents == the oracle's
    ProjectPC2SphericalRing / Voxelization, and LoadVoxelModelAndKeyPts (Match.py:46-61) reads them back."""
    from scipy import io
    from caelo_b200 import odometry
    n = odometry.preprocess_sequence(sequence["root"], frames=[0, 5], rings=True, voxels=True, keypts=True, batch=2)
    assert n == 2
    for i in (0, 5):
        pc = sequence["scans"][i]
        name = "%06d.bin.mat" % i
        m = io.loadmat(os.path.join(sequence["root"], "SphericalRing", name))
        ring, counter = oracle_mod.project_ring(pc)
        assert np.array_equal(m["SphericalRing"].view(np.uint32), ring.view(np.uint32))
        assert np.array_equal(m["GridCounter"], counter)
        v = io.loadmat(os.path.join(sequence["root"], "VoxelModel", name))
        want = oracle_mod.voxelization(pc)   # (avlBlocksList, cntVoxelsLength, AllVoxels, AllVoxels0, AllVoxels1, AllVoxels2)
        for key, w in zip(("avlBlocksList", "AllVoxels", "AllVoxels0", "AllVoxels1", "AllVoxels2"),
                          (want[0], want[2], want[3], want[4], want[5])):
            assert np.array_equal(v[key], w), key
        assert np.array_equal(v["cntVoxelsLength"].ravel(), want[1].ravel())
        kp, v0, v1, v2 = odometry.LoadVoxelModelAndKeyPts(os.path.join(sequence["raw"], "%06d.bin" % i))
        assert kp.shape == (1024, 3) and np.array_equal(v0, want[3]) and np.array_equal(v2, want[5])
        k = io.loadmat(os.path.join(sequence["root"], "KeyPts", name))
        assert k["ExtendedKeyPts"].shape[1] == 3 and k["ExtendedKeyPts"].shape[0] >= 1024
        # BatchPreprocess.py:97-105,136-141: key points are selected on the CROPPED 3-channel ring + int8 counter
        # (range gate r >= 10 m), not on the 5-channel image
        ring3, cnt8 = np.ascontiguousarray(ring[0:64, 0:1792, 0:3]), counter.astype(np.int8)
        kpo, pxo = oracle_mod.select_keypoints(ring3, cnt8, oracle_mod.respond_predict(ring3[None])[0])
        assert np.array_equal(k["KeyPts"], kpo)
        assert np.array_equal(k["ExtendedKeyPts"], oracle_mod.extend_keypoints(ring3, cnt8.copy(), pxo))
        assert np.array_equal(odometry.extended_key_points([pc])[0], k["ExtendedKeyPts"])


@pytest.mark.parametrize("n,d", [(1024, 128), (4096, 128), (16384, 128), (3000, 60)])
def test_nn_match_microbench_shapes(api, n, d):
    """configs[3]: N x N x D argmin.  Descriptors tanh(N(0,1)); frame 1 = frame-0 rows permuted, 40 % with small
    noise (known answer) and 60 % fresh.  Known answers must be hit, a random subset of columns is checked against
    float64 cdist+argmin exactly, and every index must be in range."""
    import torch
    from scipy.spatial.distance import cdist
    ctx = api.default_context()
    rng = np.random.default_rng(n + d)
    c0 = np.tanh(rng.standard_normal((n, d))).astype(np.float32)
    perm = rng.permutation(n)
    c1 = c0[perm].copy()
    noisy = rng.random(n) < 0.4
    c1[noisy] += (0.05 * rng.standard_normal((int(noisy.sum()), d))).astype(np.float32)
    c1[~noisy] = np.tanh(rng.standard_normal((int((~noisy).sum()), d))).astype(np.float32)
    if n >= 8:                       # exact duplicates: ties -> lowest row
        c1[3] = c0[5]
        c0[n - 1] = c0[5]
    got = ctx.nn_match(torch.from_numpy(c0[None]).cuda(), torch.from_numpy(c1[None]).cuda())[0].cpu().numpy()
    assert got.dtype == np.int64 and got.min() >= 0 and got.max() < n
    assert got[3] == 5
    cols = np.unique(np.r_[rng.integers(0, n, 192), 3, np.flatnonzero(noisy)[:64]])
    want = cdist(c0.astype(np.float64), c1[cols].astype(np.float64), "euclidean").argmin(axis=0)
    assert np.array_equal(got[cols], want)
    assert (got[noisy] == perm[noisy]).mean() > 0.99


# ---- f4: ICP (MyICP.py:28-73) --------------------------------------------------------------------------------
def test_nn3_transform_bit_exact(api, oracle_mod):
    import torch
    ctx = api.default_context()
    rng = np.random.default_rng(9)
    p0 = rng.uniform(-60, 60, (5000, 3)).astype(np.float32)
    p1 = np.r_[p0[:700] + rng.normal(0, 0.05, (700, 3)).astype(np.float32), rng.uniform(-60, 60, (1300, 3)).astype(np.float32)]
    p0[4000] = p0[10]                                        # duplicate: ties -> lowest index
    p1[10] = p0[10]
    idx, dist, mask, count = ctx.nn3(torch.from_numpy(p0).cuda(), torch.from_numpy(p1).cuda(), 0.1, want_mask=True)
    wi, wd = oracle_mod.nn3(p0, p1)
    assert np.array_equal(idx.cpu().numpy(), wi) and idx[10].item() == 10
    assert np.array_equal(dist.cpu().numpy(), wd)            # float64, bit for bit
    assert np.array_equal(mask.cpu().numpy().astype(bool), wd < 0.1) and int(count.item()) == int((wd < 0.1).sum())
    rt = np.array([0.99, -0.1, 0.02, 0.1, 0.99, 0.0, -0.02, 0.0, 1.0, 0.5, -0.25, 0.125], np.float32)
    d1 = torch.from_numpy(p1).cuda()
    ctx.transform_points(torch.from_numpy(rt).cuda(), d1)
    assert np.array_equal(d1.cpu().numpy(), oracle_mod.transform_points(rt[:9].reshape(3, 3), rt[9:], p1))
    i0, i1 = api.GetPtsInliners(p0, p1, 0.1)
    assert np.array_equal(i0, p0[wi[wd < 0.1]]) and np.array_equal(i1, p1[wd < 0.1])


@pytest.mark.parametrize("seq", ["00", "01"])
def test_icp_matches_oracle_and_reference_run(api, oracle_mod, seq, capsys):
    """api.ICP on the extended key points of the demo pairs == oracle.icp (same contracts: bit-identical pose,
    iteration count, inlier count, final threshold) and, within the oracle's own pinned tolerance, the UNMODIFIED
    reference run (tests/golden/icp_SS.npz)."""
    import sys
    import golden_data as G
    sys.path.insert(0, G.GOLDEN)
    import make_icp_golden as M
    z = np.load(os.path.join(G.GOLDEN, "icp_%s.npz" % seq))
    k0, k1, k1_ = M.icp_inputs(seq)
    for pc1, kw, tag in ((k1_, dict(inlierThreshold=0.3, smallShiftThreshold=0.1, ep=0.01), "tight"), (k1_, {}, "aligned"),
                         (k1, {}, "raw")):
        gi, oi = {}, {}
        R, T, ok = api.ICP(k0, pc1, info=gi, **kw)                # a batch of one through the device-side ICP
        Ro, To, oko = oracle_mod.icp(k0, pc1, info=oi, **kw)
        assert ok == oko and gi == oi, (tag, gi, oi)
        assert np.array_equal(R, Ro) and np.array_equal(T, To)
        assert R.dtype == np.float64 and T.shape == (3, 1)
        si = {}
        full = dict(maxIterTimes=50, minIterTimes=19, inlierThreshold=0.5, smallShiftThreshold=0.05, decay_rate=0.9, ep=0.001)
        full.update(kw)
        Rs, Ts, oks = api._icp_stepwise(api.default_context(), k0, pc1, info=si, **full)   # host-driven loop, brute-force search
        assert oks == ok and si == gi and np.array_equal(Rs, R) and np.array_equal(Ts, T)
        if tag == "tight":
            assert ok and np.abs(R - z["R_tight"]).max() < 1e-5 and np.abs(T - z["T_tight"]).max() < 2e-4
    assert "ICP iters:" in capsys.readouterr().out          # the reference's progress line


def test_icp_batch_equals_single_icps_and_oracle(api, oracle_mod):
    """caelo_icp_batch: seven ICPs of different sizes in ONE call (grid-indexed search, loop control on the device) ==
    the oracle's ICP pair by pair — incl. a pair that fails (< 100 inliers), one that starts far off, one with
    duplicated points (distance ties -> lowest index) and a tiny cloud."""
    import sys
    import golden_data as G
    sys.path.insert(0, G.GOLDEN)
    import make_icp_golden as M
    rng = np.random.default_rng(5)
    k0, k1, k1_ = M.icp_inputs("00")
    j0, j1, j1_ = M.icp_inputs("01")
    far = (k1_ + np.float32([0.0, 0.0, 200.0])).astype(np.float32)              # nothing within the threshold -> failure
    dup0 = np.r_[k0[:4000], k0[:4000]].astype(np.float32)                       # every target point twice
    tiny0, tiny1 = k0[:300].copy(), (k0[:300] + rng.normal(0, 0.01, (300, 3))).astype(np.float32)
    P0 = [k0, k0, j0, k0, dup0, tiny0, j0]
    P1 = [k1_, k1, j1_, far, k1_[:6000], tiny1, j1]
    kw = dict(inlierThreshold=1.0, smallShiftThreshold=0.1, ep=0.001)          # RefineOdometry's setting
    got = api.icp_batch(P0, P1, **kw)
    assert not any(g[3].get("redone_on_host") for g in got)
    for (R, T, ok, info), p0, p1 in zip(got, P0, P1):
        oi = {}
        Ro, To, oko = oracle_mod.icp(p0, p1, info=oi, **kw)
        assert ok == oko and info == oi, (info, oi)
        assert np.array_equal(R, Ro) and np.array_equal(T, To)
    assert got[3][2] is False and got[3][3]['iters'] == 1 and any(g[2] for g in got)
    one = api.icp_batch([j0], [j1_], **kw)[0]                                 # batch composition does not matter
    assert np.array_equal(one[0], got[2][0]) and np.array_equal(one[1], got[2][1]) and one[3] == got[2][3]


# ---- accuracy on a drive with known motion + the f4 refinement flow -------------------------------------------
def _ground_truth(n_frames):
    """Absolute sensor poses of synth.scan (x_world = Rz(yaw_f) x_f + pos_f) as [F,12] rows."""
    from caelo_b200 import synth
    out = []
    for f in range(n_frames):
        R, T = synth.sensor_pose(f)
        out.append(np.c_[R.astype(np.float64), T.astype(np.float64).reshape(3, 1)].reshape(12))
    return np.asarray(out)


def _pose_errors(poses, gt):
    """Per-frame relative rotation error (degrees) and translation error (m) of consecutive-frame motions — the
    RRE / RTE of EvaluationOnRegistration.py:108-130."""
    from caelo_b200 import odometry
    rre, rte = [], []
    for i in range(poses.shape[0] - 1):
        R, T = odometry.GetRelRtBetween2Poses(poses[i], poses[i + 1])
        Rg, Tg = odometry.GetRelRtBetween2Poses(gt[i], gt[i + 1])
        c = np.clip((np.trace(np.dot(Rg.T, R)) - 1) / 2, -1, 1)
        rre.append(np.degrees(np.arccos(c)))
        rte.append(np.linalg.norm(T - Tg))
    return np.asarray(rre), np.asarray(rte)


def test_odometry_and_refinement_accuracy_on_known_motion(api, sequence, capsys):
    """The whole path on a synthetic drive with known motion (0.7 m forward, 0.3 degrees yaw per frame, 2 cm range
    noise): every pair registers well inside the reference's success criterion (RRE < 1 degree, RTE < 0.5 m,
    EvaluationOnRegistration.py:23-24); the ICP refinement on the extended key points (RefinePoses.py:273-334)
    accepts every pair and does not make the trajectory worse."""
    from caelo_b200 import odometry
    scans = sequence["scans"]
    poses, rel = odometry.estimate_sequence(sequence["raw"], scans=scans, batch_pairs=8)
    assert (rel[:, 12] == 1).all()
    gt = _ground_truth(len(scans))
    rre, rte = _pose_errors(poses.astype(np.float64), gt)
    assert rre.max() < 0.5 and rte.max() < 0.2, (rre, rte)          # the paper's own figure: 0.18 deg / 0.054 +- 0.063 m
    ext = odometry.extended_key_points(scans)
    assert all(e.shape[1] == 3 and e.shape[0] > 1024 for e in ext)
    refined, codes = odometry.refine_sequence(scans, poses)
    assert (codes == 1).all() and refined.shape == poses.shape
    rre2, rte2 = _pose_errors(refined, gt)
    assert rre2.max() < 0.5 and rte2.max() < 0.2 and rte2.mean() <= rte.mean() + 0.01, (rte, rte2)
    print('RTE odometry %.3f m -> refined %.3f m; RRE %.3f -> %.3f deg' % (rte.mean(), rte2.mean(), rre.mean(), rre2.mean()))
    assert np.array_equal(refined[0], poses[0].astype(np.float64))          # the first pose is never touched
    # the batched refinement == RefinementCore + ForwardUpdatePoses pair after pair (the reference's own sequence of
    # calls, RefinePoses.py:273-334), up to the float64 rounding of re-chaining the poses F times instead of once
    from caelo_b200 import refine
    seq = poses.astype(np.float64)
    relRs, relTs = refine.all_relative_motions(seq)
    for i in range(len(scans) - 1):
        code, seq, relRs, relTs = refine.RefinementCore(seq, ext[i], ext[i + 1], i, i + 1, relRs, relTs, None)
        assert code == 1
    assert np.abs(seq - refined).max() < 1e-9
    # sharded over two 'ranks' (one-frame halo of extended key points recomputed) == single rank
    rows = [refine.refine_pairs(ext[lo:hi + 1], poses.astype(np.float64), [(i, i + 1) for i in range(lo, hi)], None, 0.5,
                                frame0=lo) for lo, hi in (pipeline_shard(len(scans) - 1, r, 2) for r in range(2))]
    assert np.array_equal(refine.chain_refined(poses, np.concatenate(rows, 0)), refined)


def pipeline_shard(n, r, w):
    from caelo_b200 import pipeline
    return pipeline.shard_pairs(n, r, w)


def test_refine_odometry_key_frames(api, sequence, tmp_path):
    """RefineOdometry (RefinePoses.py:338-475): option 0 walks consecutive pairs, option 1 the key-frame pairs found by
    transferring the inlier key points from pair to pair (GetTransferPairIdx, :102-114) — planned ahead and registered
    as one batch.  The InliersIdx files PoseEstimation.py writes (:296-309) feed the transfer, as in the reference."""
    from scipy import io
    from caelo_b200 import odometry, refine
    scans = sequence["scans"]
    idir = str(tmp_path / "InliersIdx")
    poses, rel = odometry.estimate_sequence(sequence["raw"], scans=scans, inliers_dir=idir, batch_pairs=8)
    F = len(scans)
    inl = []
    for i in range(F - 1):
        m = io.loadmat(os.path.join(idir, "%06d-%06d.bin.mat" % (i, i + 1)))
        inl.append((m["inliersIdx0"].ravel(), m["inliersIdx1"].ravel()))
    # transfer_pairs == the reference's cdist / argmin / == 0 formulation
    from scipy.spatial.distance import cdist
    a, b = inl[0][1], inl[1][0]
    D = cdist(np.c_[a, a], np.c_[b, b])
    want = [[i, int(D[i].argmin())] for i in range(D.shape[0]) if D[i, D[i].argmin()] == 0]
    assert refine.transfer_pairs(a, b) == want and len(want) > 0
    assert refine.transfer_pairs(np.zeros(0), b) == []
    ext = refine.extended_key_points(scans)
    gt = _ground_truth(F)
    p0, walk0 = refine.RefineOdometry(ext, poses, None, 0)
    assert [w[:2] for w in walk0] == [(i, i + 1) for i in range(F - 2)] and all(w[2] == 1 for w in walk0)   # :364: the last pair is never visited
    p1, walk1 = refine.RefineOdometry(ext, poses, None, 1, inliers=inl)
    assert walk1[0][0] == 0 and walk1[-1][1] >= F - 2 and all(w[2] == 1 for w in walk1)
    assert all(a[1] == b[0] for a, b in zip(walk1, walk1[1:])) and max(w[1] - w[0] for w in walk1) > 1      # real key frames
    lp = refine.longest_pair(inl, 0, F)
    assert lp == walk1[0][:2] and refine.longest_pair(inl, 0, F, nMaxTransferFrames=1) == (0, 1)
    for p in (p0, p1):
        rre, rte = _pose_errors(p, gt)
        assert rre.max() < 0.5 and rte.max() < 0.2
        assert np.array_equal(p[0], poses[0].astype(np.float64))


@pytest.mark.parametrize("seq", ["00", "01"])
def test_icp_pt2plane_matches_oracle(api, oracle_mod, seq, capsys):
    """api.ICP_Pt2PtAndPt2Plane / GetPlanarPtsInliners == the oracle restatement (bit-identical pose, counts,
    thresholds, np.random position, in-place side effect) with RefinementCore's arguments; empty planar arrays raise
    as in the reference."""
    import sys
    import golden_data as G
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_golden import _plane_inputs, PLANE_KW
    k0, k1_, pl0, pl1 = _plane_inputs(seq)
    a, b = api.GetPlanarPtsInliners(pl0, pl1[:1500], 0.5, 5.0)
    ao, bo = oracle_mod.get_planar_pts_inliers(pl0, pl1[:1500], 0.5, 5.0)
    assert np.array_equal(a, ao) and np.array_equal(b, bo) and a.shape[0] > 100
    gi, oi = {}, {}
    np.random.seed(7)
    R, T, ok = api.ICP_Pt2PtAndPt2Plane(k0, k1_, pl0, pl1.copy(), info=gi, **PLANE_KW)
    r_after = np.random.random()
    np.random.seed(7)
    Ro, To, oko = oracle_mod.icp_pt2pt_and_pt2plane(k0, k1_, pl0, pl1.copy(), info=oi, **PLANE_KW)
    assert r_after == np.random.random() and ok == oko and gi == oi, (gi, oi)
    assert np.array_equal(R, Ro) and np.array_equal(T, To)
    s1, s2 = pl1[:800].copy(), pl1[:800].copy()
    api.ICP_Pt2PtAndPt2Plane(k0, k1_, pl0, s1, **PLANE_KW)
    oracle_mod.icp_pt2pt_and_pt2plane(k0, k1_, pl0, s2, **PLANE_KW)
    assert np.array_equal(s1, s2) and not np.array_equal(s1, pl1[:800])
    assert "inliers0:" in capsys.readouterr().out
    with pytest.raises(IndexError):                    # what the shipped pipeline hands over (SphericalRing.py:219)
        api.ICP_Pt2PtAndPt2Plane(k0, k1_, np.array([], np.float32), np.array([], np.float32))
