"""Pose refinement on the extended key points (BASELINE configs[4]): what RefinePoses.py does
after the odometry, built around the batched device-side ICP (``api.icp_batch`` / ``caelo_icp_batch``).

    reference (file:line)                                       here
    LoadExtendedKeyPts + BatchPreprocess.py:136-141             extended_key_points (on the device, from raw scans)
    GetRelRtBetween2Poses / GetLidarRelRtBetween2Poses          relative_motion / lidar_relative_motion
        (Transformations.py:106-125)
    RefinementCore (RefinePoses.py:273-334)                     refine_pairs (any number of pairs per call) / RefinementCore
    ForwardUpdatePoses (RefinePoses.py:120-143)                 chain_refined (one pass) / ForwardUpdatePoses
    GetTransferPairIdx (RefinePoses.py:102-114)                 transfer_pairs
    RefineOdometry (RefinePoses.py:338-475)                     RefineOdometry (frame by frame or key frames)
    -                                                           refine_sequence: the pair list sharded over ranks

Why pairs can be refined independently (and therefore batched and sharded): RefinementCore registers frame 1's points,
moved by the CURRENT relative pose of the pair, against frame 0's.  ForwardUpdatePoses re-chains the later poses with
their stored relative motions, so refining one pair never changes the relative pose of another (beyond float64
rounding of the re-chaining).  All ICPs of a sequence can thus run side by side from the ORIGINAL relative poses; the
sequential part that is left — chaining the refined motions — is a single pass of 3x3 products on rank 0.

The shipped reference calls ICP_Pt2PtAndPt2Plane with planar points that its own pipeline never produces
(SphericalRing.py:219,285 — the call raises on the empty arrays); like round 1 this module runs the point-to-point
``ICP`` the reference keeps next to it (RefinePoses.py:297) with the thresholds of the call it replaces, and uses
``api.ICP_Pt2PtAndPt2Plane`` when a caller does supply planar points (RefinementCore only).
"""
from __future__ import annotations

import os
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import api, pipeline

ICP_KW = dict(maxIterTimes=50, minIterTimes=20 - 1, decay_rate=0.9, smallShiftThreshold=0.1, ep=0.001)   # RefinePoses.py:292-295


# ---- pose algebra ----------------------------------------------------------------------------------------------
def split_pose(pose):
    """Transformations.py:164-168 `GetRtFromOnePose`."""
    pose = np.asarray(pose).reshape(3, 4)
    return pose[:, 0:3], pose[:, 3].reshape(3, 1)


def relative_motion(pose0, pose1):
    """Transformations.py:106-113: the motion that takes pose 0 to pose 1, in the frame of the poses (camera)."""
    R0, T0 = split_pose(pose0)
    R1, T1 = split_pose(pose1)
    R0_inv = np.linalg.inv(R0)
    return np.dot(R0_inv, R1), np.dot(R0_inv, T1) - np.dot(R0_inv, T0)


def calibration(Tr):
    """(R_Tr, T_Tr, R_Tr_inv, T_Tr_inv) of the velodyne -> camera calibration row (RefinePoses.py:560-565)."""
    Tr = np.asarray(np.c_[np.eye(3), np.zeros(3)] if Tr is None else Tr, np.float32).reshape(3, 4)
    R_Tr, T_Tr = split_pose(Tr)
    R_Tr_inv = np.linalg.inv(R_Tr)
    return R_Tr, T_Tr, R_Tr_inv, -np.dot(R_Tr_inv, T_Tr)


def lidar_relative_motion(pose0, pose1, cal):
    """Transformations.py:118-125: the same motion expressed in the LiDAR frame, x0 = R x1 + T."""
    R_Tr, T_Tr, R_Tr_inv, T_Tr_inv = cal
    R0, T0 = split_pose(pose0)
    R1, T1 = split_pose(pose1)
    R0_inv = np.linalg.inv(R0)
    T0_inv = -np.dot(R0_inv, T0)
    R = np.dot(R_Tr_inv, np.dot(R0_inv, np.dot(R1, R_Tr)))
    T = np.dot(R_Tr_inv, np.dot(R0_inv, np.dot(R1, T_Tr) + T1) + T0_inv) + T_Tr_inv
    return R, T


def camera_motion(relativeR, relativeT, cal):
    """LiDAR-frame motion -> pose-frame motion (RefinePoses.py:316-317 = PoseEstimation.py:259-262)."""
    R_Tr, T_Tr, R_Tr_inv, T_Tr_inv = cal
    return np.dot(R_Tr, np.dot(relativeR, R_Tr_inv)), np.dot(R_Tr, np.dot(relativeR, T_Tr_inv) + relativeT) + T_Tr


def compose(pose0, R_diff, T_diff):
    R0, T0 = split_pose(pose0)
    return np.c_[np.dot(R0, R_diff), np.dot(R0, T_diff) + T0].reshape(12)


def ForwardUpdatePoses(poses, frameNum, newPose, relRs, relTs):
    """RefinePoses.py:120-143: replace pose ``frameNum`` and re-chain every later pose with the stored motions."""
    poses_, relRs_, relTs_ = np.array(poses, copy=True), np.array(relRs, copy=True), np.array(relTs, copy=True)
    poses_[frameNum, :] = np.asarray(newPose).reshape(12)
    R, T = relative_motion(poses_[frameNum - 1], poses_[frameNum])
    relRs_[frameNum - 1], relTs_[frameNum - 1] = R, T.reshape(3)
    for f in range(frameNum + 1, poses_.shape[0]):
        poses_[f] = compose(poses_[f - 1], relRs_[f - 1], relTs_[f - 1].reshape(3, 1))
    return poses_, relRs_, relTs_


def all_relative_motions(poses):
    F = poses.shape[0]
    relRs, relTs = np.zeros((F - 1, 3, 3), np.float64), np.zeros((F - 1, 3), np.float64)
    for i in range(F - 1):
        R, T = relative_motion(poses[i], poses[i + 1])
        relRs[i], relTs[i] = R, T.reshape(3)
    return relRs, relTs


# ---- extended key points ------------------------------------------------------------------------------------------
def extended_key_points(scans: Sequence[np.ndarray], batch: int = 16, ctx: Optional[api.Context] = None):
    """ExtendedKeyPts of every scan (BatchPreprocess.py:97-105,136-141: GetKeyPtsByAE on the cropped 3-channel ring +
    int8 counter, then ExtendKeyPtsInShpericalRing on the same arrays), batched on the device."""
    ctx = ctx or api.default_context()
    out = []
    for b0 in range(0, len(scans), batch):
        chunk = scans[b0:b0 + batch]
        off = np.zeros(len(chunk) + 1, np.int64)
        off[1:] = np.cumsum([s.shape[0] for s in chunk])
        pts = torch.from_numpy(np.ascontiguousarray(np.concatenate(chunk, 0), np.float32)).to(ctx.device)
        r = ctx.project_ring(pts, off, want=("ring3", "counter_i8"))
        _kp, px, n = ctx.select_keypoints(r["ring3"], r["counter_i8"], None, max_kpts=api.nFixedKeyPts)
        ext, n_ext = ctx.extend_keypoints(r["ring3"], r["counter_i8"], px, n)
        ne = n_ext.cpu().numpy()
        out += [ext[j, :int(ne[j])].cpu().numpy() for j in range(ext.shape[0])]
    return out


# ---- the refinement of many pairs at once --------------------------------------------------------------------------
def refine_pairs(ext, poses, pairs: Sequence[Tuple[int, int]], Tr=None, inlierThreshold0: float = 0.5, batch: int = 64,
                 ctx: Optional[api.Context] = None, frame0: int = 0):
    """RefinementCore's registration (RefinePoses.py:273-313) for every (iFrame0, iFrame1) of ``pairs`` at once:
    frame 1's extended key points moved by the pair's current relative pose (float64 product, float32 result, :285),
    batched ICP against frame 0's, the 10 degree / 5 m plausibility gate (:305-311).  ``ext[f - frame0]`` = extended
    key points of frame f.  -> rows float64 [n,16]: code (-1 ICP failed, 0 change too large, 1 refined), iFrame0,
    iFrame1, pose-frame R (9) and T (3) of the refined motion from frame 0 to frame 1, ICP iterations."""
    ctx = ctx or api.default_context()
    cal = calibration(Tr)
    rows = np.zeros((len(pairs), 16), np.float64)
    for c0 in range(0, len(pairs), batch):
        chunk = pairs[c0:c0 + batch]
        ori, moved = [], []
        for f0, f1 in chunk:
            oriRelR, oriRelT = lidar_relative_motion(poses[f0], poses[f1], cal)
            ori.append((oriRelR, oriRelT))
            moved.append(np.array((np.dot(oriRelR, ext[f1 - frame0].T) + oriRelT).T, dtype=np.float32))
        res = api.icp_batch([ext[f0 - frame0] for f0, _ in chunk], moved, inlierThreshold=inlierThreshold0, ctx=ctx, **ICP_KW)
        for k, ((f0, f1), (oriRelR, oriRelT), (R_ICP, T_ICP, ok, info)) in enumerate(zip(chunk, ori, res)):
            row = rows[c0 + k]
            row[1], row[2], row[15] = f0, f1, info["iters"]
            if not ok:
                row[0] = -1
                continue
            relativeR = np.dot(R_ICP, oriRelR)
            relativeT = np.dot(R_ICP, oriRelT) + T_ICP
            dE = np.linalg.norm(api.RotateMat2EulerAngle_XYZ(oriRelR) - api.RotateMat2EulerAngle_XYZ(relativeR))
            dT = np.linalg.norm(oriRelT - relativeT)
            if dE > 10 or dT > 5:
                row[0] = 0
                continue
            Rd, Td = camera_motion(relativeR, relativeT, cal)
            row[0] = 1
            row[3:12], row[12:15] = Rd.ravel(), Td.ravel()
    return rows


def chain_refined(poses, rows):
    """Every accepted refinement applied in ONE pass: the pose of a refined pair's frame 1 is recomputed from frame 0's
    (new) pose and the refined motion; every other pose follows from its predecessor's new pose and the ORIGINAL motion
    — what a ForwardUpdatePoses call per refined pair (RefinePoses.py:329) arrives at, in O(F) instead of O(F^2).
    Pairs must not overlap (consecutive pairs or the key-frame walk)."""
    poses = np.asarray(poses, np.float64)
    relRs, relTs = all_relative_motions(poses)
    out = poses.copy()
    refined = {int(r[2]): r for r in rows if r[0] == 1}
    for f in range(1, poses.shape[0]):
        if f in refined:
            r = refined[f]
            out[f] = compose(out[int(r[1])], r[3:12].reshape(3, 3), r[12:15].reshape(3, 1))
        else:
            out[f] = compose(out[f - 1], relRs[f - 1], relTs[f - 1].reshape(3, 1))
    return out


def RefinementCore(poses, KeyPts0, KeyPts1, iFrame0, iFrame1, relRs, relTs, Tr, inlierThreshold0=0.5,
                   PlanarPts0=None, PlanarPts1=None):
    """RefinePoses.py:273-334 for ONE pair with the reference's return values (code, poses, relRs, relTs) and its
    ForwardUpdatePoses side effect.  With planar points (N x 6) the reference's own call is made
    (ICP_Pt2PtAndPt2Plane, frame 1's planar coordinates moved by the odometry pose as well, :289-296)."""
    cal = calibration(Tr)
    if PlanarPts0 is not None and PlanarPts1 is not None and np.ndim(PlanarPts0) == 2 and np.ndim(PlanarPts1) == 2 \
            and PlanarPts0.shape[0] and PlanarPts1.shape[0]:
        oriRelR, oriRelT = lidar_relative_motion(poses[iFrame0], poses[iFrame1], cal)
        moved = np.array((np.dot(oriRelR, KeyPts1.T) + oriRelT).T, dtype=np.float32)
        planar1 = PlanarPts1.copy()
        planar1[:, 0:3] = np.array((np.dot(oriRelR, PlanarPts1[:, 0:3].T) + oriRelT).T, dtype=np.float32)
        R_ICP, T_ICP, ok = api.ICP_Pt2PtAndPt2Plane(KeyPts0, moved, PlanarPts0, planar1, maxIterTimes=50, minIterTimes=20 - 1,
                                                    inlierThreshold0=inlierThreshold0, decay_rate0=0.9, inlierThreshold1=5.0,
                                                    decay_rate1=0.9, smallShiftThreshold=0.1, ep=0.001)
        if not ok:
            return -1, np.array(poses, copy=True), relRs, relTs
        relativeR, relativeT = np.dot(R_ICP, oriRelR), np.dot(R_ICP, oriRelT) + T_ICP
        dE = np.linalg.norm(api.RotateMat2EulerAngle_XYZ(oriRelR) - api.RotateMat2EulerAngle_XYZ(relativeR))
        if dE > 10 or np.linalg.norm(oriRelT - relativeT) > 5:
            return 0, np.array(poses, copy=True), relRs, relTs
        Rd, Td = camera_motion(relativeR, relativeT, cal)
    else:
        row = refine_pairs(_Lookup({iFrame0: KeyPts0, iFrame1: KeyPts1}), poses, [(iFrame0, iFrame1)], Tr, inlierThreshold0)[0]
        if row[0] != 1:
            return int(row[0]), np.array(poses, copy=True), relRs, relTs
        Rd, Td = row[3:12].reshape(3, 3), row[12:15].reshape(3, 1)
    pose1 = compose(poses[iFrame0], Rd, Td)
    poses_, relRs, relTs = ForwardUpdatePoses(poses, iFrame1, pose1, relRs, relTs)
    return 1, poses_, relRs, relTs


class _Lookup:
    """ext[f] for a dict of frames (refine_pairs indexes ``ext[f - frame0]``)."""

    def __init__(self, d):
        self.d = d

    def __getitem__(self, f):
        return self.d[f]


# ---- key frames through inlier transfer (RefinePoses.py:102-114, 373-400) ------------------------------------------
def transfer_pairs(idx_prev, idx_next):
    """GetTransferPairIdx: for every element i of ``idx_prev`` (frame-k key-point indices that are inliers of the pair
    ending at frame k) the FIRST position j in ``idx_next`` (frame-k indices of the next pair's inliers) holding the
    same key point -> [[i, j], ...] (the reference's cdist + argmin + `== 0` test on the index values)."""
    idx_prev, idx_next = np.asarray(idx_prev).ravel(), np.asarray(idx_next).ravel()
    if idx_prev.shape[0] < 1 or idx_next.shape[0] < 1:
        return []
    first = {}
    for j, v in enumerate(idx_next.tolist()):
        first.setdefault(v, j)
    return [[i, first[v]] for i, v in enumerate(idx_prev.tolist()) if v in first]


def longest_pair(inliers, iFrame, n_poses, nMaxTransferFrames=20, nMinTransferPairs=1):
    """The key-frame pair starting at ``iFrame`` (RefinePoses.py:379-400): follow the inlier key points of pair
    (iFrame, iFrame+1) through the following pairs while at least ``nMinTransferPairs`` of them survive, for at most
    ``nMaxTransferFrames`` frames.  ``inliers[p]`` = (inliersIdx0, inliersIdx1) of pair (p, p+1)."""
    f0, f1 = iFrame, iFrame + 1
    idx1 = np.asarray(inliers[iFrame][1]).ravel()
    while idx1.shape[0] > nMinTransferPairs:
        nxt0, nxt1 = f1, f1 + 1
        if nxt1 >= n_poses - 1:
            break
        Idx0, Idx1 = (np.asarray(a).ravel() for a in inliers[nxt0])
        t = transfer_pairs(idx1, Idx0)
        if len(t) < nMinTransferPairs or f1 - f0 >= nMaxTransferFrames:
            break
        t = np.asarray(t)
        f1 = nxt1
        idx1 = Idx1[t[:, 1]]
    return f0, f1


def RefineOdometry(ext, poses, Tr=None, iOption: int = 0, inliers=None, iStartFrame: int = 0, inlierThreshold0: float = 1.0,
                   ctx: Optional[api.Context] = None, log=None):
    """RefinePoses.py:338-475: walk the sequence from ``iStartFrame``; option 0 refines every consecutive pair, option 1
    the key-frame pairs found by inlier transfer.  The reference's walk is sequential only through its failure handling
    (a failed long pair is retried as a one-frame pair, :411-431), so the walk is PLANNED assuming success, all planned
    pairs are registered in one batch, the plan is accepted up to the first failure and re-planned from there.
    -> (poses [F,12] float64, info rows [(iFrame0, iFrame1, code), ...])."""
    poses = np.asarray(poses, np.float64)
    F = poses.shape[0]
    iEnd = F - 2                                                        # :364
    done_rows, walk = [], []
    iFrame, nMax = iStartFrame, 20
    while iFrame < iEnd:
        plan, f, m = [], iFrame, nMax
        while f < iEnd:                                                 # the walk if every pair succeeds
            pair = (f, f + 1) if iOption == 0 else longest_pair(inliers, f, F, m)
            plan.append(pair)
            f, m = pair[1], 20
        rows = refine_pairs(ext, poses, plan, Tr, inlierThreshold0, ctx=ctx)
        advanced = False
        for pair, row in zip(plan, rows):
            code = int(row[0])
            walk.append((pair[0], pair[1], code))
            if log:
                log(pair, code)
            if code == 1:
                done_rows.append(row)
                iFrame, nMax, advanced = pair[1], 20, True
                continue
            if pair[1] - pair[0] > 1:                                   # :411-414 / :423-425: retry as a one-frame pair
                iFrame, nMax = pair[0], 1
            else:                                                       # :416-420: give this frame up
                iFrame, nMax = pair[0] + 1, 20
            advanced = True
            break
        else:
            iFrame = iEnd
        assert advanced or not plan
    return chain_refined(poses, done_rows), walk


# ---- a whole sequence, sharded ----------------------------------------------------------------------------------
def refine_sequence(scans: Sequence[np.ndarray], poses: np.ndarray, Tr: Optional[np.ndarray] = None,
                    inlierThreshold0: float = 0.5, rank: int = 0, world: int = 1, batch: int = 64,
                    ctx: Optional[api.Context] = None, first_frame: int = 0):
    """Frame-to-frame refinement of a whole pose file (RefineOdometry option 0 over every pair): the F-1 pairs are sharded
    contiguously over ``world`` ranks, every rank computes the extended key points of its frames and registers its
    pairs in batches, ONE gather brings the [n,16] float64 result rows to rank 0, which chains them.
    ``scans`` holds the whole sequence or — with ``first_frame`` — at least the frames this rank needs.
    -> (poses [F,12] float64, codes [F-1]) on rank 0, (None, local codes) elsewhere."""
    ctx = ctx or api.default_context()
    poses = np.asarray(poses, np.float64)
    P = poses.shape[0] - 1
    lo, hi = pipeline.shard_pairs(P, rank, world)
    ext = extended_key_points([scans[f - first_frame] for f in range(lo, hi + 1)], ctx=ctx) if hi > lo else []
    rows = refine_pairs(ext, poses, [(i, i + 1) for i in range(lo, hi)], Tr, inlierThreshold0, batch, ctx, frame0=lo)
    allrows = pipeline.gather_poses(rows, ctx.device, cap=-(-P // world), dtype=np.float64)
    if allrows is None:
        return None, rows[:, 0].astype(int)
    return chain_refined(poses, allrows), allrows[:, 0].astype(int)


# ---- bench sub-run (bench.py, BASELINE configs[4]) ------------------------------------------------------------------
def bench(ctx, pipe, data, dev, rank, world, dist, reps: int = 3):
    """ICP refinement of the step's pairs: extended key points of the P+1 frames, then all P registrations as one
    batched device ICP (odometry poses as the start).  Pairs are sharded like the odometry itself (every rank refines
    its own step), so the figure scales weakly; seq 00-10 would be 23,190 such pairs."""
    scans = data["scans"]
    F = len(scans)
    soff = np.zeros(F + 1, np.int64)
    soff[1:] = np.cumsum([s.shape[0] for s in scans])
    rel = pipe.run_device_scans(torch.from_numpy(np.concatenate(scans, 0)).to(dev), soff, None, list(range(F - 1)))
    poses = pipeline.chain_poses(rel).astype(np.float64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ext = extended_key_points(scans, ctx=ctx)
    torch.cuda.synchronize()
    t_ext = time.perf_counter() - t0
    pairs = [(i, i + 1) for i in range(F - 1)]
    refine_pairs(ext, poses, pairs, None, 1.0, 64, ctx)                  # warm-up
    ctx.profile(True)
    ctx.profile_fetch()
    secs = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = refine_pairs(ext, poses, pairs, None, 1.0, 64, ctx)
        torch.cuda.synchronize()
        secs.append(time.perf_counter() - t0)
    prof = ctx.profile_fetch()
    ctx.profile(False)
    t = torch.tensor([float(np.median(secs))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    sec = float(t.item())
    return {"what": "configs[4]: RefinePoses.py's ICP refinement (RefinementCore, threshold 1.0 m as RefineOdometry calls it) "
                    "of the step's %d pairs per GPU on their extended key points, all pairs in one batched device-side ICP "
                    "(50 iterations at most, loop control on the device), wall clock incl. the host's pose algebra and the "
                    "upload of the moved clouds" % (F - 1),
            "value": world * (F - 1) / sec, "unit": "refined frame-pairs/s", "seconds_per_batch": sec,
            "extended_key_points_per_frame": int(np.mean([e.shape[0] for e in ext])),
            "extended_key_points_seconds_per_frame": t_ext / F,
            "codes": {str(c): int((rows[:, 0] == c).sum()) for c in (-1, 0, 1)},
            "icp_iterations_mean": float(rows[:, 15].mean()),
            "device_ms_by_kernel": {k: v[1] / reps for k, v in prof.items() if k.startswith("icp_")}}
