"""GPU parity: every stage of the hot path, called through the C ABI (caelo_b200.api), against
the CPU oracle on the same inputs and against the committed golden / reference-run fixtures.

Bars: keypoint pixels, patch bits, nn-match indices, RANSAC inlier sets, trial counts and the
hypothesis/refit [R|t] are BIT-EXACT vs the oracle; descriptors within 1e-4 relative fp32
(measured: ~1e-6).  Run with ``pytest -m gpu`` on the B200 box."""
import numpy as np
import pytest

import golden_data as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from caelo_b200 import api as a
    a.default_context()
    return a


def _bits_from_f32(p):
    return (np.asarray(p).reshape(p.shape[0], -1) > 0)


# ---- a1 ------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", G.FRAMES[:2])
def test_respond_predict_bit_exact(api, oracle_mod, tag):
    f = G.frame(tag)
    model = api.load_model(api.WEIGHT_DIR + "/respond.npz")
    got = model.predict(f["ring3"][None])
    want = oracle_mod.respond_predict(f["ring3"][None])
    assert got.shape == (1, 64, 1792, 8) and got.dtype == np.float32
    assert np.array_equal(got, want)


def test_respond_predict_batch_and_odd_shape(api, oracle_mod):
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((3, 11, 37, 3)) * 20).astype(np.float32)
    x[rng.random(x.shape[:3]) < 0.3] = 0
    model = api.load_model(api.WEIGHT_DIR + "/respond.npz")
    assert np.array_equal(model.predict(x), oracle_mod.respond_predict(x))


# ---- a2 ------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", G.FRAMES)
def test_keypoints_bit_exact(api, oracle_mod, tag):
    f, rr = G.frame(tag), G.refrun(tag)
    resp = oracle_mod.respond_predict(f["ring3"][None])[0]
    # unfused: reference signature GetKeyPtsByAE(SphericalRing, GridCounter, RespondImg)
    kp5, px5, planar = api.GetKeyPtsByAE(f["ring5"], f["counter"], resp)
    assert np.array_equal(px5, rr["keypix_ring5_i32"]) and np.array_equal(kp5, rr["keypts_ring5_i32"])
    assert px5.dtype == np.int64 and kp5.dtype == np.float32 and planar.shape == (0,)
    kp3, px3, _ = api.GetKeyPtsByAE(f["ring3"], f["counter_i8"], resp)
    assert np.array_equal(px3, rr["keypix_ring3_i8"]) and np.array_equal(kp3, rr["keypts_ring3_i8"])
    # fused a1+a2: response computed in shared memory, never stored
    kpf, pxf, _ = api.GetKeyPtsFromRing(f["ring5"], f["counter"])
    assert np.array_equal(pxf, rr["keypix_ring5_i32"]) and np.array_equal(kpf, rr["keypts_ring5_i32"])
    kpf3, pxf3, _ = api.GetKeyPtsFromRing(f["ring3"], f["counter_i8"])
    assert np.array_equal(pxf3, rr["keypix_ring3_i8"])
    # the reference's own golden keypoints (set agreement)
    g = set(map(tuple, np.round(f["golden_KeyPts"].astype(np.float64), 4)))
    assert len(g & set(map(tuple, np.round(kpf3.astype(np.float64), 4)))) >= 1021


def test_keypoints_few_candidates_and_ties(api, oracle_mod):
    """< 1025 candidates (slice [-1025:-1] keeps all but the best) and exact score ties."""
    import torch
    rng = np.random.default_rng(11)
    H, W = 64, 1792
    ring = np.zeros((H, W, 3), np.float32)
    cnt = np.zeros((69, 1800), np.int32)
    rr, cc = np.meshgrid(np.arange(10, 32), np.arange(100, 140), indexing="ij")
    ring[rr, cc] = (rng.standard_normal(rr.shape + (3,)) * 3 + [20, 5, -1]).astype(np.float32)
    cnt[rr, cc] = 1
    resp = (rng.integers(0, 4, (H, W, 8))).astype(np.float32)  # few distinct values -> many exact ties
    want_k, want_p = oracle_mod.select_keypoints(ring, cnt, resp)
    ctx = api.default_context()
    kpts, kpix, n = ctx.select_keypoints(api._dev(ring[None]), api._dev(cnt[None]), api._dev(resp[None]))
    n = int(n.item())
    assert 50 < n < 1024 and n == want_p.shape[0]
    assert np.array_equal(kpix[0, :n].cpu().numpy(), want_p)
    assert np.array_equal(kpts[0, :n].cpu().numpy(), want_k)
    assert torch.count_nonzero(kpix[0, n:]).item() == 0


def test_keypoints_batched_frames(api):
    """B=4 frames in one launch == the four single-frame results."""
    ctx = api.default_context()
    ring = np.stack([G.frame(t)["ring3"] for t in G.FRAMES])
    cnt = np.stack([G.frame(t)["counter_i8"] for t in G.FRAMES])
    kpts, kpix, n = ctx.select_keypoints(api._dev(ring), api._dev(cnt), None)
    for b, t in enumerate(G.FRAMES):
        assert int(n[b].item()) == 1024
        assert np.array_equal(kpix[b].cpu().numpy(), G.refrun(t)["keypix_ring3_i8"])


# ---- a6 ------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", G.FRAMES[:2])
def test_patches_exact(api, oracle_mod, tag):
    f, rr = G.frame(tag), G.refrun(tag)
    pts, pl = api.GetPatchesList(f["golden_KeyPts"], f["vox0"], f["vox1"], f["vox2"])
    assert pts is f["golden_KeyPts"] and len(pl) == 3
    _, want, tr = oracle_mod.get_patches_list(f["golden_KeyPts"], f["vox0"], f["vox1"], f["vox2"],
                                              return_truncated=True)
    ref = G.unpack_patches(rr["patches_packed"])
    for s in range(3):
        assert pl[s].shape == (1024, 16, 16, 16, 1) and pl[s].dtype == np.float32
        assert np.array_equal(pl[s], want[s])                       # incl. the 496-NN cut rule
        neq = (_bits_from_f32(pl[s]) != _bits_from_f32(ref[s])).any(1)
        assert not (neq & ~tr[s]).any()                             # == unmodified reference run


def test_patches_float64_keypoints_and_edges(api, oracle_mod):
    f = G.frame(G.FRAMES[0])
    rng = np.random.default_rng(5)
    pts = f["golden_KeyPts"][:64].astype(np.float64) + rng.normal(0, 0.3, (64, 3))
    pts[0] = [-99.8, -99.8, -14.7]     # cube pokes outside the voxel grid (negative coordinates)
    pts[1] = [99.8, 99.8, 14.7]
    _, pl = api.GetPatchesList(pts, f["vox0"], f["vox1"], f["vox2"])
    _, want = oracle_mod.get_patches_list(pts, f["vox0"], f["vox1"], f["vox2"])
    for s in range(3):
        assert np.array_equal(pl[s], want[s])


def test_patches_too_few_voxels_raises(api):
    f = G.frame(G.FRAMES[0])
    with pytest.raises(ValueError):
        api.GetPatchesList(f["golden_KeyPts"][:4], f["vox0"][:100], f["vox1"], f["vox2"])


# ---- a3 ------------------------------------------------------------------------------------
# north_star: descriptors within 1e-4 relative fp32.  The contract tested here is ELEMENTWISE:
#     |d - d_ref| <= DESC_RTOL * |d_ref| + DESC_ATOL          for every one of the 60 components,
# and additionally the global form |d - d_ref| <= DESC_RTOL * max|d_ref|.  The absolute floor is needed by any
# two fp32 implementations (descriptor components cross zero; TF 1.14 vs torch-CPU already differ by 8.7e-7 on the
# golden files); ours is 1e-5 because the tensor cores truncate their fp32 accumulation (dense1 adds 384 MMAs
# into one accumulator): measured max |err| 8.0e-6 on the demo frames, independent of |d_ref| (profiles/r2_desc_err.txt:
# atol 5e-6 leaves 7 of 245,760 components outside, atol 1e-5 none).
DESC_RTOL = 1e-4
DESC_ATOL = 1e-5


def assert_descriptors_close(got, want):
    err = np.abs(got - want)
    assert err.max() <= DESC_RTOL * np.abs(want).max()
    bad = err > DESC_RTOL * np.abs(want) + DESC_ATOL
    assert not bad.any(), (int(bad.sum()), float(err.max()))


@pytest.mark.parametrize("tag", G.FRAMES)
def test_descriptors_vs_oracle_and_golden(api, oracle_mod, tag):
    f, rr = G.frame(tag), G.refrun(tag)
    enc = api.load_model(api.WEIGHT_DIR + "/encoder.npz")
    ref_patches = G.unpack_patches(rr["patches_packed"])       # from the unmodified reference
    feat = api.GetFeaturesFromPatches(enc, ref_patches)          # predict() boundary: float32 patches
    want = oracle_mod.get_features_from_patches(ref_patches)
    assert feat.shape == (1024, 60) and feat.dtype == np.float32
    assert_descriptors_close(feat, want)
    # golden Features/*.mat (TF 1.14): <1e-5 except rows whose patch hit a k-th-neighbour tie
    err = np.abs(feat - f["golden_Features"]).max(1)
    assert (err < 1e-5).sum() >= 1022
    # packed path (a6+a3 fused through device memory) gives the same numbers
    feat2 = api.GetFeaturesAtKeyPts(f["golden_KeyPts"], f["vox0"], f["vox1"], f["vox2"])
    _, mine = api.GetPatchesList(f["golden_KeyPts"], f["vox0"], f["vox1"], f["vox2"])
    assert np.array_equal(feat2, api.GetFeaturesFromPatches(enc, mine))


def test_encoder_dense_random_patches(api, oracle_mod):
    rng = np.random.default_rng(3)
    x = (rng.random((70, 16, 16, 16, 1)) < 0.3).astype(np.float32)   # far denser than real patches
    x[0] = 0
    x[1] = 1
    enc = api.load_model(api.WEIGHT_DIR + "/encoder.npz")
    got, want = enc.predict(x), oracle_mod.encoder_predict(x)
    assert_descriptors_close(got, want)
    assert enc.predict(x[:0]).shape == (0, 20)


def test_conv3_kernels_agree(api, oracle_mod):
    """conv3 with eight patches per MMA and dx folded into N (the default) against the one-patch kernel (M = 64,
    CAELO_CONV3_OCT=0) and the oracle, on patch counts that leave the last group of eight partly empty."""
    import os
    rng = np.random.default_rng(17)
    enc = api.load_model(api.WEIGHT_DIR + "/encoder.npz")
    for n in (1, 2, 3, 7, 8, 9, 255, 257, 600):
        x = (rng.random((n, 16, 16, 16, 1)) < rng.choice([0.002, 0.02, 0.2])).astype(np.float32)
        got = enc.predict(x)
        os.environ["CAELO_CONV3_OCT"] = "0"
        try:
            m64 = enc.predict(x)
        finally:
            del os.environ["CAELO_CONV3_OCT"]
        assert np.abs(got - m64).max() < 3e-6, n                 # one more product (A_lo W_lo), another accumulation order
        if n <= 257:
            assert_descriptors_close(got, oracle_mod.encoder_predict(x))
        assert np.array_equal(got, enc.predict(x))               # deterministic


def test_conv12_pair_kernel_matches_single_patch_kernel(api, oracle_mod):
    """conv1+conv2 with two patches per MMA and dx folded into N (the default) against the one-patch kernel
    (CAELO_CONV12_PAIR=0) and the oracle: odd and tiny patch counts, all-empty / all-full / one-voxel patches (background
    skipping in one or both patches of a pair), and the frame-ordered entry point with an even and an odd K."""
    import os
    import torch
    rng = np.random.default_rng(23)
    enc = api.load_model(api.WEIGHT_DIR + "/encoder.npz")

    def both(fn):
        got = fn()
        os.environ["CAELO_CONV12_PAIR"] = "0"
        try:
            old = fn()
        finally:
            del os.environ["CAELO_CONV12_PAIR"]
        return got, old

    for n in (1, 2, 3, 9, 255, 600):
        x = (rng.random((n, 16, 16, 16, 1)) < rng.choice([0.0005, 0.01, 0.2], size=(n, 1, 1, 1, 1))).astype(np.float32)
        x[0] = 0                                                  # an all-background patch next to whatever patch 1 is
        if n > 2:
            x[2] = 0
            x[2, 15, 0, 7, 0] = 1                                 # one voxel in a corner
        if n > 8:
            x[7] = 1
        got, old = both(lambda: enc.predict(x))
        assert np.abs(got - old).max() < 4e-6, n                  # (A_hi + A_lo)(W_hi + W_lo) vs three of the four products
        assert_descriptors_close(got, oracle_mod.encoder_predict(x))
        assert np.array_equal(got, enc.predict(x))                # deterministic
    ctx = api.default_context()
    for F, K in ((2, 6), (1, 7), (3, 1)):                        # [F,3,K] order: pairs by scale when K is even
        bits = rng.random((F, 3, K, 16, 16, 16)) < 0.02
        packed = torch.from_numpy(np.packbits(bits.reshape(F, 3, K, 128, 32), axis=-1, bitorder="little")
                                  .view(np.uint32).reshape(F, 3, K, 128).copy()).cuda()
        got, old = both(lambda: ctx.encode_frames(packed).cpu().numpy())
        assert got.shape == (F, K, 60) and np.abs(got - old).max() < 4e-6, (F, K)


def test_two_models_of_one_kind_keep_their_own_weights(api):
    """The weights live in the shared context; a model loaded later replaces them — every model puts its own back before
    it predicts (B200Model._bind), so the Keras shim may hand out several."""
    rng = np.random.default_rng(5)
    x = (rng.random((9, 16, 16, 16, 1)) < 0.02).astype(np.float32)
    a = api.load_model(api.WEIGHT_DIR + "/encoder.npz")
    ya = a.predict(x)
    w = dict(a.weights)
    w["dense_2/bias:0"] = np.asarray(w["dense_2/bias:0"], np.float32) + 0.25
    b = api.B200Model("encoder", w)
    yb = b.predict(x)
    assert np.abs(yb - ya).max() > 1e-2
    assert np.array_equal(a.predict(x), ya) and np.array_equal(b.predict(x), yb) and np.array_equal(a.predict(x), ya)


def test_encoder_rejects_non_binary(api):
    from caelo_b200._lib import CaeloError
    x = np.zeros((2, 16, 16, 16, 1), np.float32)
    x[1, 3, 3, 3, 0] = 0.5
    enc = api.load_model(api.WEIGHT_DIR + "/encoder.npz")
    with pytest.raises(CaeloError):
        enc.predict(x)


# ---- a4 ------------------------------------------------------------------------------------
def _nn(api, c0, c1):
    ctx = api.default_context()
    return ctx.nn_match(api._dev(c0[None]), api._dev(c1[None]))[0].cpu().numpy()


@pytest.mark.parametrize("seq", ["00", "01"])
def test_nn_match_golden(api, seq):
    a, b = G.PAIRS[seq]
    assert np.array_equal(_nn(api, G.frame(a)["golden_Features"], G.frame(b)["golden_Features"]),
                          G.pose(seq)["pair_idx"])            # scipy cdist + argmin (Match.py:257-258)
    u = G.usip(seq)
    assert np.array_equal(_nn(api, u["d0"], u["d1"]), u["pair_idx"])


@pytest.mark.parametrize("N,M,D", [(1, 1, 1), (5, 3, 60), (1000, 1031, 60), (2048, 2048, 128), (777, 64, 33)])
def test_nn_match_synthetic_and_ties(api, oracle_mod, N, M, D):
    rng = np.random.default_rng(N + M + D)
    c0 = np.tanh(rng.standard_normal((N, D))).astype(np.float32)
    c1 = np.tanh(rng.standard_normal((M, D))).astype(np.float32)
    if N > 4:
        c0[N // 2] = c0[1]                    # duplicate rows: ties -> lowest row index
        c1[0] = c0[1]
        c1[M // 2] = c0[N // 3] + np.float32(1e-7)
    assert np.array_equal(_nn(api, c0, c1), oracle_mod.nn_match(c0, c1))


def test_nn_match_quantised_descriptors_many_ties(api, oracle_mod):
    rng = np.random.default_rng(0)
    c0 = rng.integers(-2, 3, (600, 60)).astype(np.float32) / 4
    c1 = rng.integers(-2, 3, (500, 60)).astype(np.float32) / 4
    c0[300:] = c0[:300]
    assert np.array_equal(_nn(api, c0, c1), oracle_mod.nn_match(c0, c1))


# ---- a5 ------------------------------------------------------------------------------------
def test_solve_rt_bit_exact(api, oracle_mod):
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 4, 5, 33, 400, 1024):
        P1 = (rng.standard_normal((n, 3)) * 20).astype(np.float32)
        P0 = (P1 + rng.standard_normal((n, 3)) * 0.05 + [0.7, 0, 0]).astype(np.float32)
        R, T, c = api.SolveRT(P0, P1)
        Ro, To, co = oracle_mod.solve_rt(P0, P1)
        assert np.array_equal(R, Ro) and np.array_equal(T, To) and c == co, n
    # reflection quirk + coplanar (rank-2) input
    P1 = (rng.standard_normal((50, 3))).astype(np.float32)
    P0 = (P1 * np.float32([1, 1, -1])).astype(np.float32)
    R, T, c = api.SolveRT(P0, P1)
    Ro, To, co = oracle_mod.solve_rt(P0, P1)
    assert c == -1 and co == -1 and np.array_equal(R, Ro)
    P1[:, 2] = 0
    P0 = P1[:, [1, 0, 2]].copy()
    R, T, c = api.SolveRT(P0, P1)
    Ro, To, co = oracle_mod.solve_rt(P0, P1)
    assert np.array_equal(R, Ro) and np.array_equal(T, To) and c == co


def test_ransac_hypotheses_bit_exact(api, oracle_mod):
    a, b = G.PAIRS["00"]
    f0, f1 = G.frame(a), G.frame(b)
    pidx = G.pose("00")["pair_idx"]
    P0 = f0["golden_KeyPts"][pidx]
    P1 = f1["golden_KeyPts"]
    rng = np.random.default_rng(2)
    idx = rng.integers(0, 1024, (500, 4)).astype(np.int32)
    idx[7] = [5, 5, 9, 11]       # degenerate samples (duplicates): rank-deficient H
    idx[8] = [3, 3, 3, 3]
    idx[9] = [1, 1, 2, 2]
    ctx = api.default_context()
    res, mask, counts = ctx.ransac_round(api._dev(P0[None]), api._dev(P1[None]), None, api._dev(idx[None]),
                                         api._dev(np.array([0.4], np.float32)), None, want_counts=True)
    want_counts, want_rt = oracle_mod.ransac_score(P0, P1, idx, 0.4)
    assert np.array_equal(counts[0].cpu().numpy(), want_counts)


@pytest.mark.parametrize("seq", ["00", "01"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_solve_relative_pose(api, oracle_mod, seq, seed):
    a, b = G.PAIRS[seq]
    f0, f1, P = G.frame(a), G.frame(b), G.pose(seq)
    args = (f0["golden_KeyPts"], f0["golden_Features"], None, f1["golden_KeyPts"], f1["golden_Features"], None)
    np.random.seed(seed)
    R, T, ok, i0, i1, thr = api.SolveRelativePose(*args)
    nxt = np.random.random()
    np.random.seed(seed)
    Ro, To, oko, j0, j1, thro = oracle_mod.solve_relative_pose(*args)
    # bit-exact vs the oracle
    assert np.array_equal(R, Ro) and np.array_equal(T, To) and ok == oko and thr == thro
    assert np.array_equal(i0, j0) and np.array_equal(i1, j1)
    assert R.shape == (3, 3) and T.shape == (3, 1) and R.dtype == np.float32
    # vs the unmodified reference run (fixtures): same inliers, [R|t] within 1e-4, same RNG position
    assert np.array_equal(i0, P["idx0_%d" % seed]) and np.array_equal(i1, P["idx1_%d" % seed])
    assert np.abs(R - P["R_%d" % seed]).max() <= 1e-4
    Tr = P["T_%d" % seed].reshape(-1)
    assert np.abs(T.reshape(-1) - Tr).max() <= 1e-4 * max(1.0, np.abs(Tr).max())
    assert nxt == float(P["next_random_%d" % seed])


def test_ransac_failure_ladder(api, oracle_mod):
    rng = np.random.default_rng(0)
    P0 = (rng.standard_normal((300, 3)) * 50).astype(np.float32)
    P1 = (rng.standard_normal((300, 3)) * 50).astype(np.float32)
    np.random.seed(5)
    R, T, ok, mask, thr = api.RANSAC4RT(P0, P1, None, None)
    after = np.random.random()
    np.random.seed(5)
    Ro, To, oko, masko, thro = oracle_mod.ransac4rt(P0, P1)
    assert not ok and thr == 1.6 == thro and mask.sum() == 0 and mask.dtype == bool
    assert np.array_equal(R, np.eye(3)) and R.dtype == np.float64 and T.shape == (3, 1)
    assert after == np.random.random()


def test_ransac_small_n(api, oracle_mod):
    """N < 500: leastInliers = int(0.2 N); indices int32(u*N)."""
    rng = np.random.default_rng(4)
    P1 = (rng.standard_normal((37, 3)) * 10).astype(np.float32)
    P0 = (P1 + [1, 2, 3] + rng.standard_normal((37, 3)) * 0.01).astype(np.float32)
    P0[::3] += 5
    np.random.seed(9)
    R, T, ok, mask, thr = api.RANSAC4RT(P0, P1)
    np.random.seed(9)
    Ro, To, oko, masko, thro = oracle_mod.ransac4rt(P0, P1)
    assert ok == oko and thr == thro and np.array_equal(mask, masko)
    assert np.array_equal(np.asarray(R, np.float32), np.asarray(Ro, np.float32))
    assert np.array_equal(np.asarray(T, np.float32), np.asarray(To, np.float32))


# ---- batched pipeline (what bench.py times) -----------------------------------------------------
def test_pipeline_batch_matches_oracle(api, oracle_mod):
    """OdometryPipeline (device-resident and chunked host path) == the oracle run frame by frame and
    pair by pair with np.random.seed(pair_id) before each pair."""
    import torch
    from caelo_b200 import pipeline, synth
    d = synth.make_frames(4, seed=3)
    pair_ids = [40, 41, 42]
    pipe = pipeline.OdometryPipeline(api.default_context())
    ring_h = torch.from_numpy(d["ring3"]).pin_memory()
    cnt_h = torch.from_numpy(d["counter"]).pin_memory()
    vox_h = torch.from_numpy(d["vox"]).pin_memory()
    poses_host = pipe.run_host(ring_h, cnt_h, vox_h, d["vox_offsets"], pair_ids, chunks=3)
    smp = torch.from_numpy(pipeline.draw_samples(pair_ids, 1024, rounds=3)).cuda()
    poses_dev = pipe.run_device(ring_h.cuda(), cnt_h.cuda(), vox_h.cuda(), d["vox_offsets"], smp, pair_ids)
    assert np.array_equal(poses_host, poses_dev)
    # oracle
    off = d["vox_offsets"]
    kps, feats = [], []
    for f in range(4):
        resp = oracle_mod.respond_predict(d["ring3"][f][None])[0]
        kp, _ = oracle_mod.select_keypoints(d["ring3"][f], d["counter"][f], resp)
        v = [d["vox"][off[3 * f + s]:off[3 * f + s + 1]] for s in range(3)]
        _, pl = oracle_mod.get_patches_list(kp, *v)
        kps.append(kp)
        feats.append(oracle_mod.get_features_from_patches(pl))
    for i, pid in enumerate(pair_ids):
        # descriptors differ at the 1e-6 level between the tensor-core and the torch-CPU encoder, so the
        # match is checked on the GPU descriptors' own oracle result: same inlier count within the
        # RANSAC's sensitivity, and [R|t] within 1e-4 when the inlier sets coincide
        np.random.seed(pid)
        info = {}
        R, T, ok, i0, i1, thr = oracle_mod.solve_relative_pose(kps[i], feats[i], None, kps[i + 1], feats[i + 1], None, info)
        assert bool(poses_dev[i, 12]) == ok and abs(poses_dev[i, 14] - thr) < 1e-6
        if int(poses_dev[i, 13]) == len(i0):
            assert np.abs(poses_dev[i, :9].reshape(3, 3) - R).max() <= 1e-4
            assert np.abs(poses_dev[i, 9:12] - T.ravel()).max() <= 1e-4 * max(1.0, np.abs(T).max())
        else:
            assert abs(int(poses_dev[i, 13]) - len(i0)) <= 3


def _bench_cpu_frames(d, n_frames):
    """Oracle frame stage (respond -> select -> patches -> encoder) for frames 0..n-1 on a pool of host processes."""
    import multiprocessing as mp
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    off = d["vox_offsets"]
    jobs = [("port", d["ring3"][f], d["counter"][f], *[d["vox"][off[3 * f + s]:off[3 * f + s + 1]] for s in range(3)])
            for f in range(n_frames)]
    workers = min(os.cpu_count() or 1, 8)
    with mp.get_context("spawn").Pool(workers, initializer=bench._cpu_init, initargs=("port", 2)) as pool:
        return pool.map(bench._cpu_frame, jobs)


def test_pipeline_on_the_bench_config_is_exact_stage_by_stage(api, oracle_mod):
    """The batch bench.py times (33 synthetic frames, seed 1, pair ids 0..31) against the oracle, closing the chain
    stage by stage: key points of every frame bit-identical; descriptors of every frame inside the contract; and for
    EVERY pair the oracle's SolveRelativePose on the pipeline's own key points + descriptors with np.random.seed(pair)
    gives the pipeline's row bit for bit (R, T, success, inlier count, threshold, trial count)."""
    import torch
    from caelo_b200 import pipeline, synth
    F = 33
    d = synth.make_frames(F, seed=1, first_frame=0)
    pipe = pipeline.OdometryPipeline(api.default_context())
    pipe.keep_details = True
    ids = list(range(F - 1))
    poses = pipe.run_device(*(torch.from_numpy(d[k]).cuda() for k in ("ring3", "counter", "vox")), d["vox_offsets"], None, ids)
    det = pipe.last_details
    kp_g, ft_g = det["kpts"].cpu().numpy(), det["feat"].cpu().numpy()
    cpu = _bench_cpu_frames(d, F)
    for f in range(F):
        assert np.array_equal(kp_g[f], cpu[f][0]), f
        assert_descriptors_close(ft_g[f], cpu[f][1])
    assert (poses[:, 12] == 1).all()
    for p in ids:
        np.random.seed(p)
        info = {}
        R, T, ok, i0, i1, thr = oracle_mod.solve_relative_pose(kp_g[p], ft_g[p], None, kp_g[p + 1], ft_g[p + 1], None, info)
        assert np.array_equal(poses[p, :9], np.asarray(R, np.float32).ravel()), p
        assert np.array_equal(poses[p, 9:12], np.asarray(T, np.float32).ravel()), p
        assert bool(poses[p, 12]) == ok and int(poses[p, 13]) == len(i0) and abs(poses[p, 14] - thr) < 1e-6
        assert int(poses[p, 15]) == info["trials"]
        m = det["mask"][p].cpu().numpy().astype(bool)
        assert np.array_equal(np.flatnonzero(m), i1) and np.array_equal(det["pair_idx"][p].cpu().numpy()[m], i0)
    # the raw-scan entry (f1, f2+a6 on the device) gives the same rows
    soff = np.zeros(F + 1, np.int64)
    soff[1:] = np.cumsum([s.shape[0] for s in d["scans"]])
    poses_s = pipe.run_device_scans(torch.from_numpy(np.concatenate(d["scans"], 0)).cuda(), soff, None, ids)
    assert np.array_equal(poses_s, poses)


def test_pipeline_tail_stream_changes_nothing(api):
    """The pairs stage on its own stream (the default) and on the caller's stream give the same rows, also with several
    batches queued ahead and collected out of step; `join()` makes the caller's stream wait for the tail stream."""
    import torch
    from caelo_b200 import pipeline, synth
    d = synth.make_frames(9, seed=3)
    ring, cnt, vox = (torch.from_numpy(d[k]).cuda() for k in ("ring3", "counter", "vox"))
    voff = d["vox_offsets"]
    ids = list(range(100, 108))
    rows = {}
    for overlap in (True, False):
        pipe = pipeline.OdometryPipeline(api.default_context())
        pipe.tail_overlap = overlap
        hs = [pipe.enqueue_device(ring, cnt, vox, voff, None, ids) for _ in range(4)]
        pipe.join()
        torch.cuda.current_stream().synchronize()
        assert all(h["done"].query() for h in hs)                 # everything queued had finished when the stream drained
        rows[overlap] = [pipe.collect(h) for h in reversed(hs)]
    for a, b in zip(rows[True], rows[False]):
        assert np.array_equal(a, b)
    assert np.array_equal(rows[True][0], rows[True][-1]) and (rows[True][0][:, 12] == 1).all()


def test_pipeline_frames_with_few_keypoints(api, oracle_mod):
    """A sparse scan (fewer than 1024 candidates, SphericalRing.py:286 only asserts > 50) must not abort the batch:
    its pairs go through the per-pair path on the rows that exist — the reference's behaviour — and every other
    pair is untouched."""
    import torch
    from caelo_b200 import pipeline, synth
    d = synth.make_frames(4, seed=9)
    ring, cnt = d["ring3"].copy(), d["counter"].copy()
    ring[2, :, 110:] = 0                                   # frame 2: only 110 of 1792 columns return anything
    cnt[2, :, 110:] = 0
    pipe = pipeline.OdometryPipeline(api.default_context())
    pipe.keep_details = True
    ids = [10, 11, 12]
    full = pipe.run_device(*(torch.from_numpy(x).cuda() for x in (d["ring3"], d["counter"], d["vox"])), d["vox_offsets"], None, ids)
    np.random.seed(1234)
    before = np.random.get_state()[1].copy()
    poses = pipe.run_device(*(torch.from_numpy(x).cuda() for x in (ring, cnt, d["vox"])), d["vox_offsets"], None, ids)
    assert np.array_equal(np.random.get_state()[1], before)            # the caller's global stream is untouched
    det = pipe.last_details
    resp = oracle_mod.respond_predict(ring[2][None])[0]
    kp2, _ = oracle_mod.select_keypoints(ring[2], cnt[2], resp)
    n2 = kp2.shape[0]
    assert 50 < n2 < 1024
    assert np.array_equal(det["kpts"][2, :n2].cpu().numpy(), kp2)
    assert np.array_equal(poses[0], full[0])                             # pair (0,1) does not involve frame 2
    kp, ft = det["kpts"].cpu().numpy(), det["feat"].cpu().numpy()
    for p, (na, nb) in ((1, (1024, n2)), (2, (n2, 1024))):
        np.random.seed(ids[p])
        info = {}
        R, T, ok, i0, i1, thr = oracle_mod.solve_relative_pose(kp[p, :na], ft[p, :na], None, kp[p + 1, :nb], ft[p + 1, :nb],
                                                               None, info)
        assert bool(poses[p, 12]) == ok and int(poses[p, 13]) == len(i0) and abs(poses[p, 14] - thr) < 1e-6
        assert np.array_equal(poses[p, :9], np.asarray(R, np.float32).ravel())
        assert np.array_equal(poses[p, 9:12], np.asarray(T, np.float32).ravel())
        assert int(poses[p, 15]) == info["trials"]


def test_device_sample_stream_is_numpys(api):
    """caelo_ransac_draw_samples == int32(np.random.random(4) * N) after np.random.seed(seed) (Match.py:182-184),
    for small and 32-bit seeds, all ladder rounds, and a stream entered after failed rounds."""
    from caelo_b200 import pipeline
    ctx = api.default_context()
    seeds = [0, 1, 7, 4540, 123456789, 2 ** 31, 2 ** 32 - 1]
    for n_points in (1024, 833, 50):
        got = ctx.draw_samples(seeds, n_points, rounds=3).cpu().numpy()
        want = pipeline.draw_samples(seeds, n_points, rounds=3)
        assert got.shape == want.shape == (3, len(seeds), 500, 4)
        assert np.array_equal(got, want)
    late = ctx.draw_samples(seeds, 1024, rounds=1, rounds_done=2).cpu().numpy()
    assert np.array_equal(late[0], pipeline.draw_samples(seeds, 1024, rounds_done=2))
    # the reference's own draw, from the global stream
    np.random.seed(4540)
    first = np.array([np.int32(np.random.random((4,)) * 1024) for _ in range(3)])
    assert np.array_equal(got_rows := ctx.draw_samples([4540], 1024).cpu().numpy()[0, 0, :3], first), got_rows
    with pytest.raises(Exception):
        ctx.draw_samples([-1], 1024)
    with pytest.raises(Exception):
        ctx.draw_samples([2 ** 32], 1024)


def test_host_stream_equals_single_calls(api):
    """pipeline.run_host_stream (uploads of batch i+1 overlapping batch i, results collected one batch late) returns
    exactly what isolated run_host / run_host_scans calls return, batch by batch."""
    import torch
    from caelo_b200 import pipeline, synth
    pipe = pipeline.OdometryPipeline(api.default_context())
    d = synth.make_frames(5, seed=11)
    off = d["vox_offsets"]
    ring_h, cnt_h, vox_h = (torch.from_numpy(d[k]).pin_memory() for k in ("ring3", "counter", "vox"))
    batches, singles = [], []
    for f0, ids in ((0, [100, 101]), (2, [102, 103]), (1, [7, 8, 9])):
        f1 = f0 + len(ids) + 1
        vo = off[3 * f0:3 * f1 + 1] - off[3 * f0]
        args = (ring_h[f0:f1], cnt_h[f0:f1], vox_h[off[3 * f0]:off[3 * f1]], vo, ids)
        batches.append(("rings",) + args)
        singles.append(pipe.run_host(*args))
    soff = np.zeros(6, np.int64)
    soff[1:] = np.cumsum([s.shape[0] for s in d["scans"]])
    scans_h = torch.from_numpy(np.concatenate(d["scans"], 0)).pin_memory()
    batches.append(("scans", scans_h[soff[1]:soff[4]], soff[1:5] - soff[1], [7, 8]))
    singles.append(pipe.run_host_scans(scans_h[soff[1]:soff[4]], soff[1:5] - soff[1], [7, 8]))
    got = list(pipe.run_host_stream(iter(batches)))
    assert len(got) == len(singles)
    for g, w in zip(got, singles):
        assert np.array_equal(g, w)
    assert list(pipe.run_host_stream(iter([]))) == []


def test_bricks_two_step_equals_gather(api):
    """caelo_bricks_build + caelo_bricks_gather (and the scans variant) == caelo_gather_patches[_scans]."""
    import torch
    from caelo_b200 import synth
    ctx = api.default_context()
    d = synth.make_frames(2, seed=5)
    kpts, _px, n = ctx.select_keypoints(api._dev(d["ring3"]), api._dev(d["counter"]), None)
    vox = api._dev(d["vox"])
    want, _, _ = ctx.gather_patches(kpts, vox, d["vox_offsets"], n)
    ctx.bricks_build(vox, d["vox_offsets"])
    got = ctx.bricks_gather(kpts, n)
    assert torch.equal(got, want)
    soff = np.zeros(3, np.int64)
    soff[1:] = np.cumsum([s.shape[0] for s in d["scans"]])
    pts = api._dev(np.concatenate(d["scans"], 0))
    want_s, _, _, nvox_w, st_w = ctx.gather_patches_scans(kpts, pts, soff, n)
    nvox, st = ctx.bricks_build_scans(pts, soff)
    got_s = ctx.bricks_gather(kpts, n, nvox, st)
    assert torch.equal(got_s, want_s) and torch.equal(nvox, nvox_w) and torch.equal(st, st_w)
    assert torch.equal(got_s, want)          # Voxelization's voxel sets = the voxel lists of the synthetic frames
    with pytest.raises(Exception):           # no index for a batch of another size
        ctx.bricks_gather(kpts[:1].contiguous(), n[:1].contiguous())
