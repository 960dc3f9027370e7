"""Odometry file plumbing around the hot path: what PoseEstimation.py's
``__main__`` and the two offline batch scripts do with files, on top of the batched device pipeline.

    reference (file:line)                                   here
    LoadVoxelModelAndKeyPts (Match.py:46-61)                LoadVoxelModelAndKeyPts
    LoadKeyPtsAndFeatures (Match.py:65-72)                  LoadKeyPtsAndFeatures
    BatchProjectPC2SphericalRing (BatchPreprocess.py:44-67) preprocess_sequence(..., rings=True)
    BatchVoxelization (BatchVoxelization.py:42-64)          preprocess_sequence(..., voxels=True)
    PoseEstimation.py:185-310 (one sequence)                estimate_sequence

File formats are the reference's: ``<seq>/velodyne/NNNNNN.bin`` float32 (N,4) scans,
``<seq>/SphericalRing|VoxelModel|Features|InliersIdx/*.mat`` via scipy.io, ``poses_/SS.txt`` via
np.savetxt, ``calib/SS/calib_.txt`` (row 4 = Tr velodyne->camera).  The reference's RANSAC draws
from the unseeded global numpy stream; here pair ``i`` of a sequence uses ``np.random.seed(i)``
(the harness convention of SURVEY §8d), which makes a sharded run reproducible and rank-independent.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np
import torch
from scipy import io

from . import api, pipeline


# ---- readers (Match.py:46-72) ---------------------------------------------------------------
def LoadVoxelModelAndKeyPts(RawFileName):
    baseDir = os.path.dirname(os.path.dirname(RawFileName))
    name = RawFileName.split("/")[-1] + ".mat"
    mat = io.loadmat(os.path.join(baseDir, "VoxelModel", name))
    AllVoxels0, AllVoxels1, AllVoxels2 = mat["AllVoxels0"], mat["AllVoxels1"], mat["AllVoxels2"]
    KeyPts = io.loadmat(os.path.join(baseDir, "KeyPts", name))["KeyPts"]
    return KeyPts, AllVoxels0, AllVoxels1, AllVoxels2


def LoadKeyPtsAndFeatures(RawFileName):
    baseDir = os.path.dirname(os.path.dirname(RawFileName))
    mat = io.loadmat(os.path.join(baseDir, "Features", RawFileName.split("/")[-1] + ".mat"))
    return mat["KeyPts"], mat["Features"], mat["Weights"]


def read_scan(path: str) -> np.ndarray:
    """np.fromfile(..., float32).reshape(-1, 4) — every reference loader (e.g. BatchVoxelization.py:48)."""
    return np.fromfile(path, dtype=np.float32, count=-1).reshape([-1, 4])


def read_calib_Tr(calib_path: str) -> np.ndarray:
    """calib_.txt row 4 reshaped (3,4), float32 (PoseEstimation.py:209-210)."""
    calib = np.loadtxt(calib_path)
    return np.array(calib[4, :].reshape(3, 4), dtype=np.float32)


def list_scans(raw_dir: str) -> List[str]:
    """Frame i of a sequence is ``velodyne/%06d.bin`` (PoseEstimation.py:26,217-218)."""
    n = len([f for f in os.listdir(raw_dir) if f.endswith(".bin")])
    return [os.path.join(raw_dir, str(i).zfill(6) + ".bin") for i in range(n)]


def _stack_scans(scans: Sequence[np.ndarray]):
    off = np.zeros(len(scans) + 1, np.int64)
    off[1:] = np.cumsum([s.shape[0] for s in scans])
    host = torch.from_numpy(np.ascontiguousarray(np.concatenate(scans, 0), np.float32))
    return host, off


# ---- offline pre-stages, batched on the device ----------------------------------------------
def preprocess_sequence(seq_dir: str, frames: Optional[Sequence[int]] = None, rings: bool = True, voxels: bool = True,
                        keypts: bool = False, batch: int = 16, ctx: Optional[api.Context] = None):
    """Writes ``SphericalRing/NNNNNN.bin.mat`` {'SphericalRing','GridCounter'} (BatchPreprocess.py:64),
    ``VoxelModel/NNNNNN.bin.mat`` {'avlBlocksList','cntVoxelsLength','AllVoxels','AllVoxels0..2'}
    (BatchVoxelization.py:61) and, with ``keypts``, ``KeyPts/NNNNNN.bin.mat`` {'KeyPts','ExtendedKeyPts',
    'PlanarPts'} (BatchPreprocess.py:148) for the given frames of one sequence directory."""
    ctx = ctx or api.default_context()
    files = list_scans(os.path.join(seq_dir, "velodyne"))
    frames = list(range(len(files))) if frames is None else list(frames)
    for sub, on in (("SphericalRing", rings), ("VoxelModel", voxels), ("KeyPts", keypts)):
        if on:
            os.makedirs(os.path.join(seq_dir, sub), exist_ok=True)
    for b0 in range(0, len(frames), batch):
        ids = frames[b0:b0 + batch]
        scans = [read_scan(files[i]) for i in ids]
        host, off = _stack_scans(scans)
        pts = host.to(ctx.device)
        ring = cnt = None
        if rings or keypts:
            want = (("ring5", "counter_i32") if rings else ()) + (("ring3", "counter_i8") if keypts else ())
            r = ctx.project_ring(pts, off, want=want)
            if r["status"].any().item():
                raise IndexError("index %d is out of bounds for axis 1 with size %d" % (api.ImgW, api.ImgW))
            ring, cnt = r.get("ring5"), r.get("counter_i32")
        if voxels:
            v = ctx.voxelize(pts, off, want_blocks=True)
            if v["status"].any().item():
                raise IndexError("a point indexes outside the block grid")
            counts = v["counts"].cpu().numpy()
        if keypts:
            # BatchPreprocess.py:97-105,136-141: the key points come from the CROPPED 3-channel ring and the int8
            # counter (the range gate of SphericalRing.py:197 is then r >= 10 m, quirk 3), and so do the extended ones
            kp, px, n = ctx.select_keypoints(r["ring3"], r["counter_i8"], None, max_kpts=api.nFixedKeyPts)
            ext, n_ext = ctx.extend_keypoints(r["ring3"], r["counter_i8"], px, n)
        for j, i in enumerate(ids):
            name = os.path.basename(files[i]) + ".mat"
            if rings:
                io.savemat(os.path.join(seq_dir, "SphericalRing", name),
                           {"SphericalRing": ring[j].cpu().numpy(), "GridCounter": cnt[j].cpu().numpy()})
            if voxels:
                n0, n1, n2, nb = (int(c) for c in counts[j])
                io.savemat(os.path.join(seq_dir, "VoxelModel", name),
                           {"avlBlocksList": v["blocks"][j, :nb].cpu().numpy(),
                            "cntVoxelsLength": v["cnt"][j, :nb + 1].cpu().numpy(),
                            "AllVoxels": v["local0"][j, :n0].cpu().numpy(),
                            "AllVoxels0": v["vox"][j, 0, :n0].cpu().numpy(),
                            "AllVoxels1": v["vox"][j, 1, :n1].cpu().numpy(),
                            "AllVoxels2": v["vox"][j, 2, :n2].cpu().numpy()})
            if keypts:
                nk, ne = int(n[j].item()), int(n_ext[j].item())
                io.savemat(os.path.join(seq_dir, "KeyPts", name),
                           {"KeyPts": kp[j, :nk].cpu().numpy(), "ExtendedKeyPts": ext[j, :ne].cpu().numpy(),
                            "PlanarPts": np.array([], dtype=np.float32)})
    return len(frames)


# ---- one sequence of odometry ---------------------------------------------------------------
def estimate_sequence(raw_dir: Optional[str] = None, Tr: Optional[np.ndarray] = None, poses_path: Optional[str] = None,
                      features_dir: Optional[str] = None, inliers_dir: Optional[str] = None, batch_pairs: int = 32,
                      rank: int = 0, world: int = 1, pipe: Optional[pipeline.OdometryPipeline] = None,
                      scans: Optional[Sequence[np.ndarray]] = None, stacked=None, in_flight: int = 3,
                      stacked_first_frame: int = 0, n_frames: Optional[int] = None):
    """PoseEstimation.py:185-310 for one sequence: every consecutive frame pair -> relative [R|t] ->
    chained absolute poses -> ``poses_/SS.txt``; optionally the per-frame ``Features`` and per-pair
    ``InliersIdx`` .mat files (:283-310).  Frame pairs are sharded contiguously over ``world`` ranks
    (one-frame halo recomputed), each rank works in batches of ``batch_pairs`` pairs straight from the
    raw scans, and rank 0 gathers the pose rows — ONE collective per sequence — and runs the sequential chain.
    Input: ``raw_dir`` (velodyne/*.bin files), ``scans`` (list of (N,4) arrays) or ``stacked`` = (pts [sumN,4] f32
    torch tensor, row offsets): a pinned host tensor is streamed batch by batch (H2D of batch i+1 under the kernels
    of batch i), a CUDA tensor is used in place (``in_flight`` batches queued ahead, no host wait in between).
    ``stacked`` holds the whole sequence, or — with ``stacked_first_frame`` and ``n_frames`` (frames of the whole
    sequence) — only the frames this rank needs, starting at that frame (a rank loads its own shard).  Returns (poses [F,12] float32, rel [P,16]) on rank 0 and (None, local rel)
    elsewhere.  A rank that fails (a scan the reference would raise on) still takes part in the gather, so the
    other ranks never block; the error is raised on every rank afterwards."""
    pipe = pipe or pipeline.OdometryPipeline()
    files = None
    if stacked is not None:
        pts_all, off_all = stacked[0], np.asarray(stacked[1], np.int64)
        F = n_frames if n_frames is not None else off_all.shape[0] - 1
    elif scans is not None:
        F = len(scans)
    else:
        files = list_scans(raw_dir)
        F = len(files)
    P = F - 1
    lo, hi = pipeline.shard_pairs(P, rank, world)
    want_files = features_dir is not None or inliers_dir is not None
    pipe.keep_details = want_files
    for d in (features_dir, inliers_dir):
        if d:
            os.makedirs(d, exist_ok=True)
    rows = []
    ranges = [(b0, min(b0 + batch_pairs, hi)) for b0 in range(lo, hi, batch_pairs)]

    def batches():                                                          # read one batch ahead of the device
        for b0, b1 in ranges:
            if stacked is not None:
                j0, j1 = b0 - stacked_first_frame, b1 - stacked_first_frame
                assert j0 >= 0 and j1 + 1 < off_all.shape[0], "stacked scans do not cover this rank's frames"
                host, off = pts_all[int(off_all[j0]):int(off_all[j1 + 1])], off_all[j0:j1 + 2] - off_all[j0]
            else:
                chunk = [scans[i] if scans is not None else read_scan(files[i]) for i in range(b0, b1 + 1)]
                host, off = _stack_scans(chunk)
                host = host.pin_memory()
            yield ("scans", host, off, list(range(b0, b1)))

    def results():
        if stacked is not None and stacked[0].is_cuda:                      # whole sequence resident in HBM
            pending = []
            for _k, pts, off, ids in batches():
                pending.append(pipe.enqueue_device_scans(pts, off, None, ids))
                if len(pending) > in_flight:
                    yield pipe.collect(pending.pop(0))
            for h in pending:
                yield pipe.collect(h)
        else:
            yield from pipe.run_host_stream(batches())

    error = None
    try:
        for (b0, b1), poses_b in zip(ranges, results()):
            rows.append(poses_b)
            if want_files:
                _write_batch_files(pipe.last_details, b0, b1, lo, features_dir, inliers_dir)
    except Exception as e:                                                  # noqa: BLE001 — re-raised after the collective
        error = e
    rel_local = np.concatenate(rows, 0) if rows else np.zeros((0, 16), np.float32)
    rel, failed_ranks = pipeline.gather_poses(rel_local, pipe.dev, cap=-(-P // world), failed=error is not None,
                                              return_failed=True)
    if error is not None:
        raise error
    if failed_ranks:
        raise api._lib.CaeloError("rank(s) %s failed on their part of the sequence" % failed_ranks)
    if rel is None:
        return None, rel_local
    poses = pipeline.chain_poses(rel, Tr)
    if poses_path:
        os.makedirs(os.path.dirname(os.path.abspath(poses_path)), exist_ok=True)
        np.savetxt(poses_path, poses)
    return poses, rel


def _write_batch_files(d, b0, b1, lo, features_dir, inliers_dir):
    """Features/NNNNNN.bin.mat and InliersIdx/A-B.bin.mat of one batch (PoseEstimation.py:283-310)."""
    ids = list(range(b0, b1 + 1))                                       # frames b0..b1 -> pairs b0..b1-1
    kp, ft = d["kpts"].cpu().numpy(), d["feat"].cpu().numpy()
    pidx, mask, ok = d["pair_idx"].cpu().numpy(), d["mask"].cpu().numpy().astype(bool), d["ok"].cpu().numpy()
    if features_dir:
        first = 0 if b0 == lo else 1                                     # the halo frame was written by the previous batch
        for j in range(first, len(ids)):
            io.savemat(os.path.join(features_dir, str(ids[j]).zfill(6) + ".bin.mat"),
                       {"KeyPts": kp[j], "Features": ft[j],
                        "Weights": np.ones((kp[j].shape[0], 1), dtype=np.float32)})
    if inliers_dir:
        for j in range(b1 - b0):
            m = mask[j] if ok[j] else np.zeros_like(mask[j])
            io.savemat(os.path.join(inliers_dir, "%s-%s.bin.mat" % (str(ids[j]).zfill(6), str(ids[j + 1]).zfill(6))),
                       {"iFrame0": ids[j], "iFrame1": ids[j + 1], "inliersIdx0": pidx[j][m],
                        "inliersIdx1": np.arange(mask.shape[1])[m]})


# ---- f4: pose refinement on the extended key points lives in caelo_b200/refine.py ----------------------------------
def __getattr__(name):
    """Names that moved to ``caelo_b200.refine`` (kept importable from here)."""
    from . import refine
    moved = {"GetRtFromOnePose": "split_pose", "GetRelRtBetween2Poses": "relative_motion"}
    if name in moved:
        return getattr(refine, moved[name])
    if name in ("ForwardUpdatePoses", "extended_key_points", "RefinementCore", "refine_sequence", "RefineOdometry"):
        return getattr(refine, name)
    if name == "GetLidarRelRtBetween2Poses":
        return lambda pose0, pose1, R_Tr, T_Tr, R_Tr_inv, T_Tr_inv: refine.lidar_relative_motion(
            pose0, pose1, (R_Tr, T_Tr, R_Tr_inv, T_Tr_inv))
    raise AttributeError(name)
