"""TEST INFRASTRUCTURE ONLY — CPU oracle for the CAE-LO hot path.

Restates, on the CPU, what the reference computes between "ring image / voxel lists in
memory" and "relative [R|t] known".  The C half (``caelo_oracle.c``) carries the
arithmetic contracts the CUDA kernels are compared with bit-for-bit; the Python half
restates the integer patch gather in numpy and the Keras encoder graph in torch-CPU fp32
(Keras/TensorFlow are not installable here).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  Nothing under ``caelo_b200/`` does.

Pinning status (see tests/test_oracle_golden.py and tests/golden/README.md):
  a1+a2  keypoints   pinned: DemoData Features/*.mat KeyPts (set agreement; quirk 2 drift)
                     + bit-identical to the imported reference GetKeyPtsByAE
  a6+a3  descriptors pinned: DemoData Features/*.mat Features (<1e-5 on non-truncated patches)
  a4     nn match    pinned: bit-identical to scipy cdist+argmin (the reference's own call)
  a5     RANSAC      parity UNPINNED by reference data (RANSAC is unseeded there); pinned
                     against the imported reference run with harness seeds (fixtures).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcaelo_oracle.so")
_WEIGHT_DIR = os.path.join(os.path.dirname(_HERE), "caelo_b200", "weights")

# ---- constants, derived exactly as the reference derives them -------------------------
# Voxel.py:15-52
VoxelSize = 0.02
PatchSize = 16
BlockRealSize = 1.28
ScaleRatios = [1, 8, 32]
VoxelSizes = [VoxelSize, VoxelSize * ScaleRatios[1], VoxelSize * ScaleRatios[2]]
PatchRadius = int(PatchSize / 2)
nBlocksL = int(2 * 100 / BlockRealSize)
nBlocksW = int(2 * 100 / BlockRealSize)
nBlocksH = int(2 * 15 / BlockRealSize)
VisibleLength = nBlocksL / 2 * BlockRealSize
VisibleWidth = nBlocksW / 2 * BlockRealSize
VisibleHeight = nBlocksH / 2 * BlockRealSize
# SphericalRing.py:28-57
nLines = 64
ImgH = 69
ImgW = 1800
CropWidth_SphericalRing = 8
N_NEIGHBORS = 496  # Voxel.py:182


def build(force: bool = False) -> str:
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "caelo_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_select.restype = ctypes.c_int
        _lib.oracle_kabsch.restype = ctypes.c_int
        _lib.oracle_ransac_replay.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---- weights ------------------------------------------------------------------------------
def load_weights(which: str):
    """``which`` in {"respond", "encoder"} -> dict of float32 arrays (exported blobs)."""
    z = np.load(os.path.join(_WEIGHT_DIR, which + ".npz"))
    return {k: z[k] for k in z.files}


# ---- a1 -----------------------------------------------------------------------------------
def respond_predict(x: np.ndarray, w=None) -> np.ndarray:
    """RespondLayer.predict restated (contract R1).  x: (B,H,W,3) f32 -> (B,H,W,8) f32."""
    w = w or load_weights("respond")
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, H, W, C = x.shape
    assert C == 3
    out = np.empty((B, H, W, 8), np.float32)
    w1 = np.ascontiguousarray(w["conv2d_1/kernel:0"], np.float32)
    b1 = np.ascontiguousarray(w["conv2d_1/bias:0"], np.float32)
    w2 = np.ascontiguousarray(w["conv2d_2/kernel:0"].reshape(32, 8), np.float32)
    b2 = np.ascontiguousarray(w["conv2d_2/bias:0"], np.float32)
    lib().oracle_respond(_p(x), B, H, W, _p(w1), _p(b1), _p(w2), _p(b2), _p(out))
    return out


# ---- a2 -----------------------------------------------------------------------------------
def select_keypoints(ring: np.ndarray, counter: np.ndarray, resp: np.ndarray, maxk: int = 1024,
                     return_score: bool = False):
    """GetKeyPtsByAE restated (contract S1).  Returns (KeyPts f32 (n,3), KeyPixels i64 (n,2))."""
    ring = np.ascontiguousarray(ring, np.float32)
    resp = np.ascontiguousarray(resp, np.float32)
    H, W, C8 = resp.shape
    assert C8 == 8
    if counter.dtype == np.int8:
        kind = 0
    else:
        counter = counter.astype(np.int32, copy=False)
        kind = 1
    counter = np.ascontiguousarray(counter)
    kpts = np.zeros((maxk, 3), np.float32)
    kpix = np.zeros((maxk, 2), np.int64)
    score = np.zeros((H, W), np.float32) if return_score else None
    n = lib().oracle_select(_p(resp), H, W, _p(ring), ring.shape[2], ring.shape[0], ring.shape[1],
                            _p(counter), kind, counter.shape[0], counter.shape[1], maxk,
                            _p(kpts), _p(kpix), _p(score) if return_score else None)
    if return_score:
        return kpts[:n].copy(), kpix[:n].copy(), score
    return kpts[:n].copy(), kpix[:n].copy()



# ---- f1 / f2: the offline pre-stages --------------------------------------------
import math as _math

AzimuthResolution = 0.20 * (_math.pi / 180)                       # SphericalRing.py:34,48
VerticalViewDown = -24.8 * (_math.pi / 180)
VerticalViewUp = 2.0 * (_math.pi / 180)
VerticalResolution = (VerticalViewUp - VerticalViewDown) / (nLines - 1)
VerticalPixelsOffset = -VerticalViewDown / VerticalResolution


def project_ring(PC: np.ndarray):
    """ProjectPC2SphericalRing restated (contract P1, SphericalRing.py:72-94).
    PC (N,4) f32 -> (ring (69,1800,5) f32, counter (69,1800) i32); raises IndexError where numpy does."""
    pc = np.ascontiguousarray(PC, np.float32)
    assert pc.shape[0] > 3 and pc.shape[1] == 4
    ring = np.zeros((ImgH, ImgW, 5), np.float32)
    counter = np.zeros((ImgH, ImgW), np.int32)
    f = lib().oracle_project_ring
    f.restype = ctypes.c_int
    bad = f(_p(pc), ctypes.c_int64(pc.shape[0]), ImgH, ImgW, ctypes.c_double(AzimuthResolution),
            ctypes.c_double(VerticalResolution), ctypes.c_double(VerticalPixelsOffset), _p(ring), _p(counter))
    if bad:
        raise IndexError("index %d is out of bounds for axis 1 with size %d" % (ImgW, ImgW))
    return ring, counter


def voxelization(PC: np.ndarray):
    """Voxelization restated (contract V1, Voxel.py:89-173) -> (avlBlocksList, cntVoxelsLength,
    AllVoxels, AllVoxels0, AllVoxels1, AllVoxels2) exactly as BatchVoxelization.py:61 stores them."""
    pc = np.ascontiguousarray(PC, np.float32)
    N, C = pc.shape
    v0, v1, v2, loc = (np.zeros((N, 3), np.int16) for _ in range(4))
    blocks = np.zeros((N, 3), np.int16)
    cnt = np.zeros((N + 1,), np.int32)
    counts = np.zeros((4,), np.int32)
    f = lib().oracle_voxelize
    f.restype = ctypes.c_int
    rc = f(_p(pc), ctypes.c_int64(N), C, _p(v0), _p(v1), _p(v2), _p(loc), _p(blocks), _p(cnt), _p(counts))
    if rc:
        raise IndexError("a point indexes outside the block grid (the reference raises here too)")
    n0, n1, n2, nb = (int(c) for c in counts)
    return (blocks[:nb].copy(), cnt[:nb + 1].copy(), loc[:n0].copy(), v0[:n0].copy(), v1[:n1].copy(), v2[:n2].copy())


def extend_keypoints(SphericalRing, GridCounter, KeyPixels):
    """ExtendKeyPtsInShpericalRing restated (SphericalRing.py:294-317): every still-occupied pixel of each key
    pixel's 13x13 window, in key-pixel order then row-major; the counter window is zeroed IN PLACE."""
    out = []
    for iX, iY in np.asarray(KeyPixels).reshape(-1, 2):
        oneMask = GridCounter[iX - 6:iX + 7, iY - 6:iY + 7]
        out.append(SphericalRing[iX - 6:iX + 7, iY - 6:iY + 7, 0:3][oneMask > 0])
        oneMask[:] = 0
    return np.concatenate(out, 0).astype(np.float32) if out else np.zeros((0, 3), np.float32)

# ---- a6 -----------------------------------------------------------------------------------
def _pack(v):
    v = v.astype(np.int64)
    return (v[..., 0] << 40) | (v[..., 1] << 20) | v[..., 2]


_BALL = None


def _ball_offsets():
    global _BALL
    if _BALL is None:
        r = np.arange(-13, 14)
        g = np.stack(np.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3)
        d2 = (g * g).sum(1)
        g = g[d2 <= 3 * PatchRadius * PatchRadius]
        _BALL = g
    return _BALL


def get_patches_list(Pts, AllVoxels0, AllVoxels1, AllVoxels2, return_truncated: bool = False):
    """GetPatchesList restated (Voxel.py:177-216).  Integer-exact.

    The reference asks sklearn for the 496 nearest occupied voxels of each key voxel and
    keeps those inside the [-8,8)^3 cube; the scatter uses negative indices, so offset o is
    stored at index o mod 16.  Here a k-d tree (scipy) returns the same 496 neighbours; when
    the 496th lies beyond the cube's farthest corner (d^2 > 192) the cut cannot bite and the
    result is order-independent.  Otherwise ("truncated") the set is decided by the oracle's
    canonical rule: a voxel survives iff its rank among all occupied voxels ordered by
    (d^2, x, y, z) is < 496 — the tie order at the k-th neighbour is implementation-defined
    in sklearn, parity unpinned there.
    """
    from scipy.spatial import cKDTree

    Pts = np.asarray(Pts)
    K = Pts.shape[0]
    Pts_ = Pts + [VisibleLength, VisibleWidth, VisibleHeight]  # float64, as the reference
    patches = []
    truncated = []
    ball = _ball_offsets()
    R2 = 3 * PatchRadius * PatchRadius
    for s, vox in enumerate((AllVoxels0, AllVoxels1, AllVoxels2)):
        vox = np.asarray(vox)
        if vox.shape[0] < N_NEIGHBORS:
            raise ValueError("Expected n_neighbors <= n_samples_fit")  # sklearn's own error
        KeyVoxels = np.array(Pts_ / VoxelSizes[s], dtype=np.int32)
        out = np.zeros((K, PatchSize, PatchSize, PatchSize, 1), np.float32)
        trunc = np.zeros(K, bool)
        if K == 0:
            patches.append(out)
            truncated.append(trunc)
            continue
        vox64 = vox.astype(np.int64)
        tree = cKDTree(vox64.astype(np.float64))
        _, nbr = tree.query(KeyVoxels.astype(np.float64), k=N_NEIGHBORS)
        off = vox64[nbr] - KeyVoxels[:, None, :].astype(np.int64)       # (K,496,3)
        d2 = (off * off).sum(-1)
        trunc = d2.max(1) <= R2                                          # the cut can bite
        incube = ((off >= -PatchRadius) & (off < PatchRadius)).all(-1)
        kk, nn = np.nonzero(incube & ~trunc[:, None])
        o = off[kk, nn]
        out[kk, o[:, 0] % 16, o[:, 1] % 16, o[:, 2] % 16, 0] = 1.0
        if trunc.any():
            keys = np.unique(_pack(vox))
            for i in np.flatnonzero(trunc):
                kv = KeyVoxels[i].astype(np.int64)
                b = kv + ball
                bvalid = (b >= 0).all(-1)
                q = _pack(np.where(bvalid[:, None], b, 0))
                pos = np.minimum(np.searchsorted(keys, q), keys.size - 1)
                hit = (keys[pos] == q) & bvalid
                bo = ball[hit]
                bc = kv + bo
                order = np.lexsort((bc[:, 2], bc[:, 1], bc[:, 0], (bo * bo).sum(1)))
                keep = bo[order[:N_NEIGHBORS]]
                keep = keep[((keep >= -PatchRadius) & (keep < PatchRadius)).all(1)]
                out[i, keep[:, 0] % 16, keep[:, 1] % 16, keep[:, 2] % 16, 0] = 1.0
        patches.append(out)
        truncated.append(trunc)
    if return_truncated:
        return Pts, patches, truncated
    return Pts, patches


# ---- a3 -----------------------------------------------------------------------------------
def encoder_predict(patches: np.ndarray, w=None, batch_size: int = 32) -> np.ndarray:
    """PatchEncoder.predict restated in torch-CPU fp32 (AE4VoxelPatch.py:189-197 topology,
    activations per the shipped .h5 = tanh everywhere)."""
    import torch
    import torch.nn.functional as F

    w = w or load_weights("encoder")
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items()}
    k1 = t["conv3d_1/kernel:0"].permute(4, 3, 0, 1, 2).contiguous()
    k2 = t["conv3d_2/kernel:0"].permute(4, 3, 0, 1, 2).contiguous()
    k3 = t["conv3d_3/kernel:0"].permute(4, 3, 0, 1, 2).contiguous()
    outs = []
    x_all = torch.from_numpy(np.ascontiguousarray(patches, np.float32))
    with torch.no_grad():
        for i in range(0, x_all.shape[0], batch_size):  # Keras predict default batch_size=32
            x = x_all[i:i + batch_size].permute(0, 4, 1, 2, 3)
            x = torch.tanh(F.conv3d(x, k1, t["conv3d_1/bias:0"], padding=1))
            x = F.max_pool3d(x, 2)
            x = torch.tanh(F.conv3d(x, k2, t["conv3d_2/bias:0"], padding=1))
            x = F.max_pool3d(x, 2)
            x = torch.tanh(F.conv3d(x, k3, t["conv3d_3/bias:0"], padding=1))
            x = x.permute(0, 2, 3, 4, 1).reshape(x.shape[0], -1)  # channels-last Flatten
            x = torch.tanh(x @ t["dense_1/kernel:0"] + t["dense_1/bias:0"])
            x = torch.tanh(x @ t["dense_2/kernel:0"] + t["dense_2/bias:0"])
            outs.append(x.numpy())
    if not outs:
        return np.zeros((0, 20), np.float32)
    return np.concatenate(outs, 0)


def get_features_from_patches(patches_list, w=None) -> np.ndarray:
    """GetFeaturesFromPatches (Match.py:130-135)."""
    return np.c_[tuple(encoder_predict(p, w) for p in patches_list)]


# ---- a4 -----------------------------------------------------------------------------------
def nn_match(codes0: np.ndarray, codes1: np.ndarray, return_dist: bool = False):
    """cdist(...,'euclidean') + argmin(axis=0) restated (contract M1)."""
    c0 = np.ascontiguousarray(codes0, np.float32)
    c1 = np.ascontiguousarray(codes1, np.float32)
    N, D = c0.shape
    M = c1.shape[0]
    idx = np.zeros(M, np.int64)
    dist = np.zeros(M, np.float64)
    lib().oracle_nn_match(_p(c0), N, _p(c1), M, D, _p(idx), _p(dist))
    return (idx, dist) if return_dist else idx


# ---- a5 -----------------------------------------------------------------------------------
def solve_rt(P0: np.ndarray, P1: np.ndarray):
    """SolveRT (Match.py:138-158) under contract K1 -> (R f32 (3,3), T f32 (3,1), isCredible)."""
    p0 = np.ascontiguousarray(P0, np.float32)
    p1 = np.ascontiguousarray(P1, np.float32)
    R = np.zeros(9, np.float32)
    T = np.zeros(3, np.float32)
    cred = ctypes.c_int(0)
    n = lib().oracle_kabsch(_p(p0), _p(p1), None, None, p0.shape[0], _p(R), _p(T), ctypes.byref(cred))
    if n <= 0:
        raise ValueError("SolveRT needs at least one point")
    return R.reshape(3, 3), T.reshape(3, 1), cred.value


def solve_rt_masked(P0: np.ndarray, P1: np.ndarray, mask: np.ndarray):
    """SolveRT over the rows with mask!=0, summed in ORIGINAL-index lanes (contract K1) — the
    order the device refit uses; differs from solve_rt(P0[mask], P1[mask]) only in the last
    float64 bit of the sums."""
    p0 = np.ascontiguousarray(P0, np.float32)
    p1 = np.ascontiguousarray(P1, np.float32)
    m = np.ascontiguousarray(mask, np.uint8)
    R = np.zeros(9, np.float32)
    T = np.zeros(3, np.float32)
    cred = ctypes.c_int(0)
    n = lib().oracle_kabsch(_p(p0), _p(p1), _p(m), None, p0.shape[0], _p(R), _p(T), ctypes.byref(cred))
    if n <= 0:
        raise ValueError("SolveRT needs at least one point")
    return R.reshape(3, 3), T.reshape(3, 1), cred.value


def ransac_score(P0, P1, sample_idx, thr):
    p0 = np.ascontiguousarray(P0, np.float32)
    p1 = np.ascontiguousarray(P1, np.float32)
    si = np.ascontiguousarray(sample_idx, np.int32)
    Tn = si.shape[0]
    counts = np.zeros(Tn, np.int32)
    Rt = np.zeros((Tn, 12), np.float32)
    lib().oracle_ransac_score(_p(p0), _p(p1), p0.shape[0], _p(si), Tn, ctypes.c_float(thr),
                              _p(counts), _p(Rt))
    return counts, Rt


MAX_TRIALS = 500


def ransac4rt(Pairs0, Pairs1, Weights0=None, Weights1=None, info=None):
    """RANSAC4RT (Match.py:162-218) restated; consumes the global ``np.random`` stream exactly
    as the reference does (4 doubles per trial) by drawing a full round,
    replaying the sequential rule, then rewinding and re-drawing what was really used."""
    p0 = np.ascontiguousarray(Pairs0, np.float32)
    p1 = np.ascontiguousarray(Pairs1, np.float32)
    N = p0.shape[0]
    thr = 0.4
    best_n = 0
    R_star = np.eye(3, dtype=np.float64)
    T_star = np.zeros((3, 1), dtype=np.float64)
    mask_star = np.zeros((N,), dtype=bool)
    ok = False
    total_trials = 0
    while True:
        state = np.random.get_state()
        u = np.random.random((MAX_TRIALS, 4))
        idx = np.array(u * N, dtype=np.int32)
        counts, Rt = ransac_score(p0, p1, idx, thr)
        bt, bn, okc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        used = lib().oracle_ransac_replay(_p(counts), MAX_TRIALS, N, best_n, ctypes.byref(bt),
                                          ctypes.byref(bn), ctypes.byref(okc))
        np.random.set_state(state)
        np.random.random((used * 4,))
        total_trials = used
        if bt.value >= 0:
            best_n = bn.value
            R_star = Rt[bt.value, :9].reshape(3, 3).copy()
            T_star = Rt[bt.value, 9:].reshape(3, 1).copy()
            m = np.zeros(N, np.uint8)
            lib().oracle_inlier_mask(_p(p0), _p(p1), N, _p(np.ascontiguousarray(Rt[bt.value])),
                                     ctypes.c_float(thr), _p(m))
            mask_star = m.astype(bool)
        if okc.value:
            ok = True
            break
        thr = 2 * thr
        if thr > 2.0:
            thr = thr / 2
            break
    if info is not None:
        info.update(trials=total_trials, n_inliers=int(best_n))
    return R_star, T_star, ok, mask_star, thr


def solve_relative_pose(PC0, Codes0, W0, PC1, Codes1, W1, info=None):
    """SolveRelativePose (Match.py:241-283) restated."""
    PC0 = np.asarray(PC0)
    PC1 = np.asarray(PC1)
    pair_idx = nn_match(Codes0, Codes1)
    Pairs0 = PC0[pair_idx, :]
    R, T, ok, mask, thr = ransac4rt(Pairs0, PC1, None, None, info)
    idx0 = pair_idx[mask]
    idx1 = np.arange(PC1.shape[0])[mask]
    if idx0.shape[0] == 0:
        return R, T, ok, idx0, idx1, thr
    R, T, _ = solve_rt_masked(Pairs0, PC1, mask)  # == SolveRT(PC0[idx0], PC1[idx1]) (Match.py:280-282)
    return R, T, ok, idx0, idx1, thr


# ---- f4: ICP on the extended key points (MyICP.py:28-73, 76-85, 127-201) ------------------------------------
RADIAN2DEGREE = 180.0 / np.pi   # Transformations.py


def rotate_mat_to_euler_xyz(R):
    """RotateMat2EulerAngle_XYZ (Transformations.py:181-186), degrees, float64."""
    import math
    a = np.zeros((3,))
    a[0] = math.atan2(R[2, 1], R[2, 2]) * RADIAN2DEGREE
    a[1] = math.atan2(-R[2, 0], math.sqrt(math.pow(R[2, 1], 2) + math.pow(R[2, 2], 2))) * RADIAN2DEGREE
    a[2] = math.atan2(R[1, 0], R[0, 0]) * RADIAN2DEGREE
    return a


def nn3(PC0: np.ndarray, PC1: np.ndarray):
    """Exact 1-nearest neighbour of every PC1 point among PC0 (what sklearn's kd-tree returns for
    NearestNeighbors(n_neighbors=1), MyICP.py:33-34), contract N1: float32 inputs widened to float64,
    d = sqrt(((dx*dx) + (dy*dy)) + (dz*dz)); ties -> lowest PC0 index.  -> (idx int64 [M], dist float64 [M])."""
    from scipy.spatial import cKDTree
    p0 = np.ascontiguousarray(PC0, np.float32).astype(np.float64)
    p1 = np.ascontiguousarray(PC1, np.float32).astype(np.float64)
    k = min(4, p0.shape[0])
    _, cand = cKDTree(p0).query(p1, k=k)                     # candidates; the contract arithmetic decides among them
    cand = cand.reshape(p1.shape[0], k)
    d = p0[cand] - p1[:, None, :]
    dk = np.sqrt(((d[..., 0] * d[..., 0]) + (d[..., 1] * d[..., 1])) + (d[..., 2] * d[..., 2]))
    best = dk.min(axis=1, keepdims=True)
    idx = np.where(dk == best, cand, np.iinfo(np.int64).max).min(axis=1)   # exact ties -> lowest index
    dist = best[:, 0]
    return idx.astype(np.int64), dist


def transform_points(R, T, PC):
    """PC1 = (R PC1^T + T)^T (MyICP.py:51) under contract U1: float64 products and sums in the order
    ((r0*x + r1*y) + r2*z) + t, one rounding to float32."""
    R = np.asarray(R, np.float32).astype(np.float64).reshape(3, 3)
    T = np.asarray(T, np.float32).astype(np.float64).reshape(3)
    p = np.ascontiguousarray(PC, np.float32).astype(np.float64)
    out = np.empty(p.shape, np.float64)
    for a in range(3):
        out[:, a] = ((R[a, 0] * p[:, 0] + R[a, 1] * p[:, 1]) + R[a, 2] * p[:, 2]) + T[a]
    return out.astype(np.float32)


def icp(PC0, PC1, maxIterTimes=50, minIterTimes=20 - 1, inlierThreshold=0.5, smallShiftThreshold=0.05,
        decay_rate=0.9, ep=0.001, min_inliers=100, decay_rate1=None, info=None):
    """ICP (MyICP.py:28-73) restated: per iteration exact 1-NN (N1), inliers dist < threshold, SolveRT on them
    (K1, summed in original-index lanes like the device refit), PC1 update (U1), R*/T* accumulation in float64,
    Euler-angle / translation convergence test and threshold decay exactly as the reference's loop.
    -> (R_star f64 (3,3), T_star f64 (3,1), isSuccess)."""
    R_star = np.eye(3, dtype=np.float64)
    T_star = np.zeros((3, 1), dtype=np.float64)
    PC0 = np.ascontiguousarray(PC0, np.float32)
    PC1 = np.ascontiguousarray(PC1, np.float32)
    n_in, it_done = 0, 0
    for iIter in range(maxIterTimes):
        idx, dist = nn3(PC0, PC1)
        mask = dist < inlierThreshold
        n_in = int(mask.sum())
        it_done = iIter + 1
        if n_in < min_inliers:
            if info is not None:
                info.update(iters=it_done, inliers=n_in, threshold=inlierThreshold)
            return R_star, T_star, False
        R, T, _ = solve_rt_masked(PC0[idx], PC1, mask)
        PC1 = transform_points(R, T, PC1)
        R_star = np.dot(R, R_star)
        T_star = np.dot(R, T_star) + T
        normEulers = np.linalg.norm(rotate_mat_to_euler_xyz(R))
        normT = np.linalg.norm(T)
        if iIter >= minIterTimes:
            if normEulers < ep and normT < ep:
                break
        if normEulers < smallShiftThreshold and normT < smallShiftThreshold:
            inlierThreshold *= decay_rate
    if info is not None:
        info.update(iters=it_done, inliers=n_in, threshold=inlierThreshold)
    return R_star, T_star, True


def get_pts_inliers(PC0, PC1, inlierThreshold):
    """GetPtsInliners (MyICP.py:76-85)."""
    idx, dist = nn3(PC0, PC1)
    m = dist < inlierThreshold
    return PC0[idx[m], :], PC1[m, :]


def planar_pedals(inliers0, inliers1, norms1):
    """The point-to-plane part of GetPlanarPtsInliners (MyICP.py:103-109) exactly as numpy evaluates it on the
    input dtype: foot points of frame-0 points on the frame-1 tangent planes, and their distance to the plane point."""
    vetors = inliers0 - inliers1
    dist2Planes = np.sum(norms1 * vetors, axis=1)
    pedals = inliers1 + norms1 * np.tile(dist2Planes.reshape(dist2Planes.shape[0], 1), [1, 3])
    distances = np.linalg.norm((pedals - inliers1), axis=1)
    return pedals, distances


def get_planar_pts_inliers(PtsWithNorm0, PtsWithNorm1, inlierThreshold0, inlierThreshold1):
    """GetPlanarPtsInliners (MyICP.py:88-113)."""
    PC0, PC1, Norms1 = PtsWithNorm0[:, 0:3], PtsWithNorm1[:, 0:3], PtsWithNorm1[:, 3:6]
    idx, dist = nn3(PC0, PC1)
    m = dist < inlierThreshold1
    inliers0, inliers1 = PC0[idx[m], :], PC1[m, :]
    pedals, distances = planar_pedals(inliers0, inliers1, Norms1[m, :])
    keep = (distances < inlierThreshold0).flatten()
    return pedals[keep, :], inliers1[keep, :]


def icp_pt2pt_and_pt2plane(PC0, PC1, PtsWithNorm0, PtsWithNorm1, maxIterTimes=50, minIterTimes=20 - 1,
                           inlierThreshold0=0.5, decay_rate0=0.9, inlierThreshold1=2.0, decay_rate1=0.5,
                           smallShiftThreshold=0.1, ep=0.01, info=None):
    """ICP_Pt2PtAndPt2Plane (MyICP.py:127-201) restated: point pairs + planar foot-point pairs stacked into one
    SolveRT per iteration; more than 2000 planar points are subsampled from the global np.random stream (:135-139);
    PtsWithNorm1's coordinates are updated IN PLACE (its normals are not rotated), as the reference does."""
    R_star = np.eye(3, dtype=np.float64)
    T_star = np.zeros((3, 1), dtype=np.float64)
    PC0 = np.ascontiguousarray(PC0, np.float32)
    PC1 = np.ascontiguousarray(PC1, np.float32)
    nMaxPts = 2000
    if PtsWithNorm1.shape[0] > nMaxPts:
        RandIdxes = np.random.random((nMaxPts,))
        RandIdxes = RandIdxes * (PtsWithNorm1.shape[0])
        RandIdxes = np.array(RandIdxes, dtype=np.int32)
        PtsWithNorm1 = PtsWithNorm1[RandIdxes, :]
    isSuccess = True
    n_pts = n_pl = 0
    it_done = 0
    for iIter in range(maxIterTimes):
        it_done = iIter + 1
        in0_pts, in1_pts = get_pts_inliers(PC0, PC1, inlierThreshold0)
        in0_pl, in1_pl = get_planar_pts_inliers(PtsWithNorm0, PtsWithNorm1, inlierThreshold0, inlierThreshold1)
        n_pts, n_pl = in0_pts.shape[0], in0_pl.shape[0]
        inliers0 = np.r_[in0_pts, in0_pl]
        inliers1 = np.r_[in1_pts, in1_pl]
        if inliers0.shape[0] < 200:
            if iIter < 1:
                isSuccess = False
            break
        R, T, _ = solve_rt(inliers0, inliers1)
        PC1 = transform_points(R, T, PC1)
        PtsWithNorm1[:, 0:3] = transform_points(R, T, PtsWithNorm1[:, 0:3])
        R_star = np.dot(R, R_star)
        T_star = np.dot(R, T_star) + T
        normEulers = np.linalg.norm(rotate_mat_to_euler_xyz(R))
        normT = np.linalg.norm(T)
        if iIter >= minIterTimes:
            if normEulers < ep and normT < ep:
                break
        if normEulers < smallShiftThreshold and normT < smallShiftThreshold:
            inlierThreshold0 *= decay_rate0
            inlierThreshold1 *= decay_rate1
    if info is not None:
        info.update(iters=it_done, inliers0=n_pts, inliers1=n_pl, th0=inlierThreshold0, th1=inlierThreshold1)
    return R_star, T_star, isSuccess
