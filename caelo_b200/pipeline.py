"""Batched odometry pipeline: F consecutive frames -> F-1 relative poses on one GPU, and the
frame-pair sharding across ranks.

This is what replaces the reference's producer/consumer fan-out in PoseEstimation.py:48-150,
241-267 (four loader processes + one Keras process + a CPU RANSAC loop): every stage is batched
over frames (a1/a2/a6/a3) or frame pairs (a4/a5) and runs as a handful of kernel launches.
Per pair the arithmetic is exactly that of ``api.SolveRelativePose`` with ``np.random.seed(pair_id)``
called before it (the harness convention of SURVEY §8d)."""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import api

MAX_TRIALS = api.MAX_TRIALS


def shard_pairs(n_pairs: int, rank: int, world: int):
    """Contiguous pair range [lo, hi) of this rank; it needs frames [lo, hi] (one-frame halo
    recomputed, not exchanged)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


LADDER = (0.4, 0.8, 1.6)   # residualThreshold of the three RANSAC4RT rounds (Match.py:170,209-211)


def draw_samples(pair_ids: Sequence[int], n_points: int, rounds_done: int = 0, rounds: int = 1) -> np.ndarray:
    """Sample indices of RANSAC round ``rounds_done`` ([P,500,4] int32), or of ``rounds`` consecutive
    rounds ([rounds,P,500,4]), drawn exactly as RANSAC4RT does after ``np.random.seed(pair_id)``
    (Match.py:182-184: 4 doubles per trial, int32(u*N); a failed round consumes all 500 trials)."""
    out = np.empty((rounds, len(pair_ids), MAX_TRIALS, 4), np.int32)
    rs = np.random.RandomState(0)                      # one object, re-seeded per pair (constructing one costs ~0.3 ms)
    for i, pid in enumerate(pair_ids):
        rs.seed(int(pid))
        if rounds_done:
            rs.random_sample((rounds_done * MAX_TRIALS * 4,))
        u = rs.random_sample((rounds, MAX_TRIALS, 4))
        out[:, i] = np.array(u * n_points, dtype=np.int32)
    return out[0] if rounds == 1 else out


class OdometryPipeline:
    def __init__(self, ctx: Optional[api.Context] = None, n_keypoints: int = api.nFixedKeyPts):
        self.ctx = ctx or api.default_context()
        self.K = n_keypoints
        self.dev = self.ctx.device
        self.keep_details = False      # True: keep the batch's keypoints / descriptors / matches / inlier masks
        self.last_details = None       # (device tensors) for the Features / InliersIdx files of odometry.py
        self.tail_overlap = os.environ.get("CAELO_TAIL_STREAM", "1") != "0"
        self.last_done = None          # event of the most recent batch's pairs stage

    # ---- device-resident stages ---------------------------------------------------------
    def frames_to_descriptors(self, ring: torch.Tensor, counter: torch.Tensor, vox: torch.Tensor,
                              vox_offsets: np.ndarray):
        """ring [F,64,1792,3] (or [F,69,1800,5]), counter [F,69,1800] i8/i32, voxel lists ->
        kpts [F,K,3], feat [F,K,60], n_kpts [F]."""
        # (building the a6 index on a second stream under the selection was measured: the two compete for the same SMs,
        #  brick_insert 0.19 -> 0.45 ms, respond_score 0.40 -> 0.60 ms, step 4.69 -> 4.75 ms — so one stream)
        kpts, _kpix, n = self.ctx.select_keypoints(ring, counter, None, max_kpts=self.K)
        packed, _, _ = self.ctx.gather_patches(kpts, vox, vox_offsets, n, reuse=True)
        feat = self.ctx.encode_frames(packed)
        return kpts, feat, n

    def scans_to_descriptors(self, pts: torch.Tensor, pts_offsets: np.ndarray):
        """Raw scans (pts [sumN,4] f32 + host row offsets [F+1]) -> kpts, feat, n_kpts, status [F]:
        f1 projection -> a1+a2 -> f2+a6 fused (bricks straight from the points) -> a3.  ``status`` is
        non-zero for a frame the reference would have raised on (IndexError / sklearn ValueError)."""
        r = self.ctx.project_ring(pts, pts_offsets, want=("ring3", "counter_i8"), reuse=True)
        kpts, _kpix, n = self.ctx.select_keypoints(r["ring3"], r["counter_i8"], None, max_kpts=self.K)
        packed, _, _, _nvox, st = self.ctx.gather_patches_scans(kpts, pts, pts_offsets, n, reuse=True)
        feat = self.ctx.encode_frames(packed)
        return kpts, feat, n, st | r["status"]

    # ---- whole batch ------------------------------------------------------------------------
    def _enqueue_pairs(self, kpts, feat, n, samples, pair_ids, status=None):
        """Queues the pairs stage (a4 + a5 + refit) and the one D2H copy of the per-pair results on the current
        stream; nothing here waits for the device.  ``samples`` is [P,500,4] (first round; failures go through a
        host-driven ladder in ``_collect``) or [3,P,500,4]: then the 0.8 and 1.6 rounds are queued on the device
        right away and skip every pair that already has a model (no host round trip)."""
        pc0, pc1 = kpts[:-1], kpts[1:]
        rounds = (samples if samples.dim() == 4 else samples[None]).contiguous()
        # The pairs stage is ~25 small latency-bound launches (0.3 ms per 32-pair batch at < 10 % of the SMs' issue slots).
        # It goes on its own stream behind an event, so that the NEXT batch's frame stages — queued on the caller's stream
        # right after this returns — run next to it instead of behind it.  Its scratch (misc, match_ops, pose_ws, seed_ws)
        # is disjoint from the frame stages' (cand, bricks, scan_ws, enc_ws), its inputs are this batch's own tensors, and
        # consecutive batches' pairs stages stay ordered on the one tail stream.  CAELO_TAIL_STREAM=0: everything on the
        # caller's stream.
        cur = torch.cuda.current_stream(self.dev)
        tail = self._tail_stream_() if self.tail_overlap else cur
        if tail is not cur:
            ready = torch.cuda.Event()
            ready.record(cur)
            tail.wait_event(ready)
            for t in (kpts, feat, n, rounds) + (() if status is None else (status,)):
                t.record_stream(tail)                   # allocated under the caller's stream, read under the tail stream
        with torch.cuda.stream(tail):
            pair_idx = self.ctx.nn_match(feat[:-1], feat[1:])
            state, mask_acc, rt, thr_used = self.ctx.ransac_ladder(pc0, pc1, pair_idx, rounds, LADDER[:rounds.shape[0]])
            details = None
            if self.keep_details:
                details = dict(kpts=kpts, feat=feat, pair_idx=pair_idx, mask=mask_acc, ok=state[:, 12] != 0)
            nf = n.to(torch.float32)
            bad = torch.zeros_like(nf) if status is None else (status != 0).to(torch.float32)
            packed = torch.cat([state, rt, thr_used[:, None], nf[:-1, None], nf[1:, None],
                                torch.maximum(bad[:-1], bad[1:])[:, None]], 1)
            host = self._result_slot(packed.shape)
            host.copy_(packed, non_blocking=True)
            done = torch.cuda.Event()
            done.record(tail)
        self.last_done = done
        return dict(host=host, done=done, kpts=kpts, feat=feat, pair_idx=pair_idx, pair_ids=list(pair_ids),
                    rounds=rounds.shape[0], details=details)

    def _tail_stream_(self):
        if not hasattr(self, "_tail_stream"):
            self._tail_stream = torch.cuda.Stream(self.dev)
        return self._tail_stream

    def join(self):
        """Makes the caller's stream wait for every pairs stage queued so far (they run on the tail stream): call before
        recording an event that is meant to cover whole batches."""
        if self.last_done is not None:
            torch.cuda.current_stream(self.dev).wait_event(self.last_done)

    def _result_slot(self, shape):
        """Pinned host buffer for one batch's result rows, from a free list that ``_collect`` refills — no pinned
        allocation on the hot path once as many buffers exist as batches are ever in flight."""
        if not hasattr(self, "_res_free"):
            self._res_free = []
        for i, buf in enumerate(self._res_free):
            if tuple(buf.shape) == tuple(shape):
                return self._res_free.pop(i)
        return torch.empty(tuple(shape), dtype=torch.float32, pin_memory=True)

    def _collect(self, h):
        """Waits for one queued batch (the one sync point of the batch); returns poses [P,16] float32 (host):
        refit R(9) T(3), isSuccess, nInliers, residualThreshold, trials — rows follow ``pair_ids``."""
        h["done"].synchronize()
        if self.keep_details:
            self.last_details = h["details"]
        host = h["host"].numpy().copy()                 # the pinned slot is reused by a later batch
        self._res_free.append(h["host"])
        P = host.shape[0]
        res, rt_h = host[:, :16], host[:, 16:28]
        if host[:, 31].any():
            raise api._lib.CaeloError("a scan has points outside the voxel grid / ring image or fewer than 496 "
                                      "occupied voxels at some scale (the reference raises on it too)")
        poses = np.zeros((P, 16), np.float32)
        poses[:, :12] = rt_h
        poses[:, 12] = res[:, 12]
        poses[:, 13] = res[:, 14]
        poses[:, 14] = host[:, 28]
        poses[:, 15] = res[:, 13]
        failed = np.flatnonzero(res[:, 12] == 0)
        short = np.flatnonzero((host[:, 29] != self.K) | (host[:, 30] != self.K))
        # the rare host-driven fallbacks use the pairs stage's scratch: they run where the pairs stages run
        with torch.cuda.stream(self._tail_stream_() if self.tail_overlap else torch.cuda.current_stream(self.dev)):
            if failed.size and h["rounds"] < len(LADDER):
                self._ladder(failed, h["kpts"], h["pair_idx"], h["pair_ids"], poses, h["rounds"])
            elif failed.size:                               # total failure: R=I, T=0 (Match.py:277-278)
                poses[failed, :12] = [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]
                poses[failed, 13] = 0
            if short.size:
                self._short_pairs(short, host[:, 29].astype(int), host[:, 30].astype(int), h, poses)
        return poses

    def _short_pairs(self, short, n0s, n1s, h, poses):
        """Pairs with a frame that yielded fewer than K key points (a sparse scan): the batched kernels assume K
        rows, so these pairs are redone through the per-pair path on the rows that exist — exactly what the
        reference does with whatever GetKeyPtsByAE returned (it only asserts n > 50, SphericalRing.py:286) — with
        np.random.seed(pair_id) as everywhere in the pipeline; the caller's global stream is left untouched."""
        state = np.random.get_state()
        try:
            for p in short:
                n0, n1 = int(n0s[p]), int(n1s[p])
                assert n0 > 50 and n1 > 50                              # SphericalRing.py:286
                kp0, kp1 = h["kpts"][p, :n0].cpu().numpy(), h["kpts"][p + 1, :n1].cpu().numpy()
                ft0, ft1 = h["feat"][p, :n0].cpu().numpy(), h["feat"][p + 1, :n1].cpu().numpy()
                np.random.seed(int(h["pair_ids"][p]))
                R, T, ok, i0, _i1, thr, used = api.solve_relative_pose(self.ctx, kp0, ft0, kp1, ft1)
                poses[p, :9] = np.asarray(R, np.float32).ravel()
                poses[p, 9:12] = np.asarray(T, np.float32).ravel()
                poses[p, 12:] = [1.0 if ok else 0.0, len(i0), thr, used]
        finally:
            np.random.set_state(state)

    def enqueue_device(self, ring, counter, vox, vox_offsets, samples, pair_ids):
        """Queues one batch whose inputs are already in HBM and returns a handle for ``collect`` — nothing waits for
        the device, so consecutive batches run back to back.  ``samples`` None: the RANSAC indices of all three
        ladder rounds are generated on the device from the pair ids (np.random.seed(pair_id) streams)."""
        kpts, feat, n = self.frames_to_descriptors(ring, counter, vox, vox_offsets)
        if samples is None:
            samples = self.ctx.draw_samples(pair_ids, self.K, rounds=3)
        return self._enqueue_pairs(kpts, feat, n, samples, pair_ids)

    def enqueue_device_scans(self, pts, pts_offsets, samples, pair_ids):
        """The same from raw scans already in HBM."""
        kpts, feat, n, st = self.scans_to_descriptors(pts, pts_offsets)
        if samples is None:
            samples = self.ctx.draw_samples(pair_ids, self.K, rounds=3)
        return self._enqueue_pairs(kpts, feat, n, samples, pair_ids, st)

    def collect(self, handle):
        """Waits for a queued batch -> poses [P,16] float32 (host)."""
        return self._collect(handle)

    def run_device(self, ring, counter, vox, vox_offsets, samples, pair_ids):
        """Inputs already in HBM: enqueue_device + collect."""
        return self._collect(self.enqueue_device(ring, counter, vox, vox_offsets, samples, pair_ids))

    def run_device_scans(self, pts, pts_offsets, samples, pair_ids):
        """Raw scans already in HBM."""
        return self._collect(self.enqueue_device_scans(pts, pts_offsets, samples, pair_ids))

    # ---- host (pinned) inputs -----------------------------------------------------------------
    def _copy_stream_(self):
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.dev)
        return self._copy_stream

    def _slot_buffer(self, slot: dict, name: str, like: torch.Tensor, rows: int) -> torch.Tensor:
        """Device staging buffer ``name`` of an upload slot with at least ``rows`` rows shaped/typed like ``like``
        (grown geometrically, never shrunk): the hot path allocates nothing for its inputs."""
        buf = slot.get(name)
        if buf is None or buf.shape[0] < rows or buf.shape[1:] != like.shape[1:] or buf.dtype != like.dtype:
            if buf is not None:
                torch.cuda.current_stream(self.dev).synchronize()        # rare: a bigger batch than ever before
                self._copy_stream_().synchronize()
            buf = torch.empty((int(rows * 1.25) + 1,) + tuple(like.shape[1:]), dtype=like.dtype, device=self.dev)
            slot[name] = buf
        return buf

    def _upload(self, batch, chunks: int = 4):
        """Queues the H2D copies of one batch on the copy stream in ``chunks`` frame groups, into one of two
        preallocated device slots (a slot is reused once the kernels of the batch that last used it have been
        queued AND have run: its `free` event); returns the device parts with the event each becomes valid at.
        ``batch`` is ("rings", ring_h, counter_h, vox_h, vox_offsets, pair_ids) or ("scans", pts_h, pts_offsets,
        pair_ids), host tensors pinned."""
        cs = self._copy_stream_()
        if not hasattr(self, "_slots"):
            # THREE upload slots: with two, the copy of batch i+1 can only start once batch i-1's frame stages have run,
            # i.e. it lands on the latency-bound match / RANSAC tail of batch i-1 (those kernels were measured 1.6-3.4x
            # slower next to the copy, +0.46 ms per step); with three the slot of batch i+1 has been free since batch
            # i-2, the copy starts the moment the host queues it — at the beginning of batch i-1's tensor-core phase —
            # and is over before the tail begins.  CAELO_UPLOAD_SLOTS=2 restores the old behaviour for A/B timing.
            n_slots = max(2, int(os.environ.get("CAELO_UPLOAD_SLOTS", "3")))
            self._slots, self._slot_next = [dict() for _ in range(n_slots)], 0
        slot = self._slots[self._slot_next]
        self._slot_next = (self._slot_next + 1) % len(self._slots)
        if "free" in slot:
            cs.wait_event(slot["free"])
        parts = []
        if batch[0] == "rings":
            _, ring_h, counter_h, vox_h, vox_offsets, pair_ids = batch
            F = ring_h.shape[0]
            voff = np.asarray(vox_offsets, np.int64)
            d_ring = self._slot_buffer(slot, "ring", ring_h, F)
            d_cnt = self._slot_buffer(slot, "cnt_" + str(counter_h.dtype), counter_h, F)
            d_vox = self._slot_buffer(slot, "vox", vox_h, int(voff[-1]))
            bounds = np.linspace(0, F, min(chunks, F) + 1).astype(int)
            for c0, c1 in zip(bounds[:-1], bounds[1:]):
                v0, v1 = int(voff[3 * c0]), int(voff[3 * c1])
                with torch.cuda.stream(cs):
                    d_ring[c0:c1].copy_(ring_h[c0:c1], non_blocking=True)
                    d_cnt[c0:c1].copy_(counter_h[c0:c1], non_blocking=True)
                    d_vox[v0:v1].copy_(vox_h[v0:v1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                parts.append(((d_ring[c0:c1], d_cnt[c0:c1], d_vox[v0:v1]), voff[3 * c0:3 * c1 + 1] - voff[3 * c0], ev))
        else:
            _, pts_h, pts_offsets, pair_ids = batch
            off = np.asarray(pts_offsets, np.int64)
            F = off.shape[0] - 1
            d_pts = self._slot_buffer(slot, "pts", pts_h, int(off[-1]))
            bounds = np.linspace(0, F, min(chunks, F) + 1).astype(int)
            for c0, c1 in zip(bounds[:-1], bounds[1:]):
                p0, p1 = int(off[c0]), int(off[c1])
                with torch.cuda.stream(cs):
                    d_pts[p0:p1].copy_(pts_h[p0:p1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                parts.append(((d_pts[p0:p1],), off[c0:c1 + 1] - off[c0], ev))
        return dict(kind=batch[0], parts=parts, pair_ids=list(pair_ids), slot=slot)

    def _enqueue(self, up):
        """Queues every kernel of an uploaded batch on the current stream (each frame group as soon as its copy
        event has fired); the RANSAC sample indices are generated on the device (ctx.draw_samples)."""
        cur = torch.cuda.current_stream(self.dev)
        outs = []
        for tensors, o, ev in up["parts"]:
            cur.wait_event(ev)
            if up["kind"] == "rings":
                kp, ft, n = self.frames_to_descriptors(*tensors, o)
                outs.append((kp, ft, n, None))
            else:
                outs.append(self.scans_to_descriptors(tensors[0], o))
        smp = self.ctx.draw_samples(up["pair_ids"], self.K, rounds=3)     # all three ladder rounds, on the device
        kpts, feat, n = (torch.cat([o[i] for o in outs], 0) for i in range(3))
        st = None if up["kind"] == "rings" else torch.cat([o[3] for o in outs], 0)
        # The slot is released as soon as every kernel that reads the inputs is queued (the frame stages), not at the
        # end of the batch: on this box the 75 MB upload takes ~4.6 ms (16 GB/s) — longer than the 4.2 ms of kernels —
        # so it must start as early as possible.  (Releasing at the end of the batch keeps the copy away from the
        # latency-bound match / RANSAC tail, whose kernels run up to 2x slower next to it, but then the device waits
        # for the copy: 4.91 vs 4.64 ms per step measured.)
        free = torch.cuda.Event()
        free.record(cur)
        up["slot"]["free"] = free
        return self._enqueue_pairs(kpts, feat, n, smp, up["pair_ids"], st)

    def run_host_stream(self, batches, chunks: int = 1, first_chunks: int = 4):
        """Generator over an iterable of host batches (see ``_upload``) -> poses [P,16] per batch, in order.
        Software-pipelined two deep: the H2D copies of batch i+1 run on the copy stream and its kernels are
        queued while batch i computes; the host only waits for batch i-1's results.  Every batch still pays
        its own H2D copies and D2H read — they just overlap the neighbouring batches' kernels (this is how
        ``odometry.estimate_sequence`` walks a sequence).  Only the first batch is uploaded in ``first_chunks``
        frame groups (nothing else is running that could hide its copy); the others go up whole, so that every
        kernel sees the full batch (per-frame stages such as the top-k are latency-bound per launch)."""
        it = iter(batches)
        try:
            up = self._upload(next(it), first_chunks)
        except StopIteration:
            return
        pending = None
        while up is not None:
            try:
                nxt = self._upload(next(it), chunks)
            except StopIteration:
                nxt = None
            h = self._enqueue(up)
            if pending is not None:
                yield self._collect(pending)
            pending, up = h, nxt
        yield self._collect(pending)

    def run_host_scans(self, pts_h: torch.Tensor, pts_offsets: np.ndarray, pair_ids: Sequence[int], chunks: int = 4):
        """End-to-end from raw scans in HOST (pinned) memory: [sumN,4] f32 + row offsets [F+1].  Same
        chunked upload / compute overlap as ``run_host``."""
        self._copy_stream_().wait_stream(torch.cuda.current_stream(self.dev))
        return self._collect(self._enqueue(self._upload(("scans", pts_h, pts_offsets, pair_ids), chunks)))

    def run_host(self, ring_h: torch.Tensor, counter_h: torch.Tensor, vox_h: torch.Tensor,
                 vox_offsets: np.ndarray, pair_ids: Sequence[int], chunks: int = 4):
        """End-to-end call with HOST (pinned) buffers.  Frames are uploaded in ``chunks`` groups on a copy
        stream while the compute stream already works on the groups that have landed; the RANSAC sample
        indices are drawn on the host while the GPU is busy with the frame stages."""
        self._copy_stream_().wait_stream(torch.cuda.current_stream(self.dev))
        return self._collect(self._enqueue(self._upload(("rings", ring_h, counter_h, vox_h, vox_offsets, pair_ids), chunks)))

    def _ladder(self, failed, kpts, pair_idx, pair_ids, poses, rounds_done=1):
        """Host-driven threshold ladder for the pairs that still have no model (Match.py:207-214)."""
        thr_val = LADDER[rounds_done - 1]
        rounds = rounds_done - 1
        while failed.size:
            thr_val *= 2
            rounds += 1
            if thr_val > 2.0:
                for p in failed:                        # total failure: R=I, T=0 (Match.py:277-278)
                    poses[p, :12] = [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]
                    poses[p, 12:] = [0, 0, thr_val / 2, MAX_TRIALS]
                return
            sel = torch.as_tensor(failed, device=self.dev)
            pc0 = kpts[:-1][sel].contiguous()
            pc1 = kpts[1:][sel].contiguous()
            pidx = pair_idx[sel].contiguous()
            smp = torch.from_numpy(draw_samples([pair_ids[p] for p in failed], self.K, rounds)).to(self.dev)
            thr = torch.full((failed.size,), thr_val, dtype=torch.float32, device=self.dev)
            result, mask, _ = self.ctx.ransac_round(pc0, pc1, pidx, smp, thr)
            rt, _ = self.ctx.kabsch(pc0, pc1, pidx, mask)
            res, rt_h = result.cpu().numpy(), rt.cpu().numpy()
            still = []
            for i, p in enumerate(failed):
                if res[i, 12] != 0:
                    poses[p, :12] = rt_h[i]
                    poses[p, 12:] = [1, res[i, 14], thr_val, res[i, 13]]
                else:
                    still.append(p)
            failed = np.asarray(still, np.int64)


_comm_streams = {}


def gather_poses(poses: np.ndarray, device: torch.device, cap: Optional[int] = None, failed: bool = False,
                 return_failed: bool = False, dtype=np.float32):
    """NCCL gather of the per-rank [P_local,16] pose rows to rank 0 (the only collective on the
    path; PoseEstimation.py:254-267's pose chain then runs on rank 0).  Returns the concatenated
    array on rank 0, None elsewhere.  Works without torch.distributed initialised (1 rank).
    ``cap``: an upper bound of P_local that every rank knows (e.g. ceil(P_total / world)); then ONE gather of
    [cap+1,16] rows moves everything (row 0 carries the row count) instead of a count exchange first.  The
    collective runs on its own CUDA stream, so it never waits for kernels queued on the compute stream.
    ``failed``: this rank hit an error on its part — it still takes part (so nobody blocks) and flags it in row 0;
    with ``return_failed`` the call returns (rows, [failed ranks]) (the list is only known on rank 0).
    ``dtype``: float32 pose rows of the odometry, float64 for the refinement's rows (caelo_b200.refine)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return (poses, [0] if failed else []) if return_failed else poses
    world, rank = dist.get_world_size(), dist.get_rank()
    import contextlib
    if device.type == "cuda":
        comm = _comm_streams.get(device)
        if comm is None:
            comm = _comm_streams[device] = torch.cuda.Stream(device)
        on_comm = torch.cuda.stream(comm)
    else:                                   # gloo (the CPU tests of the sharding logic)
        on_comm = contextlib.nullcontext()
    rows = np.ascontiguousarray(poses, dtype)
    with on_comm:
        if cap is None:
            n_loc = torch.tensor([rows.shape[0]], dtype=torch.int64, device=device)
            counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
            dist.all_gather(counts, n_loc)
            cap = max(int(c.item()) for c in counts)
        assert rows.shape[0] <= cap
        buf = np.zeros((cap + 1, 16), dtype)
        buf[0, 0] = rows.shape[0]
        buf[0, 1] = 1.0 if failed else 0.0
        buf[1:1 + rows.shape[0]] = rows
        t = torch.from_numpy(buf).to(device)
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, bufs, dst=0)
        if rank != 0:
            return (None, []) if return_failed else None
        allr = torch.stack(bufs).cpu().numpy()          # waits for the comm stream only
    out = np.concatenate([allr[r, 1:1 + int(allr[r, 0, 0])] for r in range(world)], 0)
    bad = [r for r in range(world) if allr[r, 0, 1] != 0]
    return (out, bad) if return_failed else out


def chain_poses(rel: np.ndarray, Tr: Optional[np.ndarray] = None):
    """The sequential pose chain of PoseEstimation.py:209-267 on rank 0: rel [P,16] rows (R(9) T(3)
    isSuccess ...) -> absolute poses [P+1,12] float32 (KITTI format, what np.savetxt writes at :272).
    Tr is the 3x4 velodyne->camera calibration row (calib_.txt line 5, :209-214); identity if None.
    Dtypes follow the reference: everything float32 until a pair fails — its R = I, T = 0 are float64
    (Match.py:277-278) and promote the rest of the chain."""
    Tr = np.asarray(np.c_[np.eye(3), np.zeros(3)] if Tr is None else Tr, dtype=np.float32).reshape(3, 4)
    R_Tr = Tr[:, 0:3]
    R_Tr_inv = np.linalg.inv(R_Tr)
    T_Tr = Tr[:, 3].reshape(3, 1)
    T_Tr_inv = -np.dot(R_Tr_inv, T_Tr)
    # the recurrence is inherently sequential and its float32 np.dot calls are kept exactly as the reference makes them
    # (a batched np.matmul rounds differently); only the Python overhead around them is trimmed — for seq 00 this loop
    # is what rank 0 does alone after the one gather
    P = rel.shape[0]
    ok = rel[:, 12] != 0
    out = np.empty((P + 1, 3, 4), dtype=np.float32)
    R0 = np.eye(3, dtype=np.float32)
    T0 = np.zeros((3, 1), dtype=np.float32)
    out[0, :, :3], out[0, :, 3:] = R0, T0
    rel32 = np.ascontiguousarray(rel[:, :12], dtype=np.float32)
    rel64 = rel32.astype(np.float64) if not ok.all() else None
    dot = np.dot
    for i in range(P):
        src = rel32[i] if ok[i] else rel64[i]
        relativeR = src[:9].reshape(3, 3)
        relativeT = src[9:12].reshape(3, 1)
        R_poseDiff = dot(R_Tr, dot(relativeR, R_Tr_inv))
        T_poseDiff = dot(R_Tr, dot(relativeR, T_Tr_inv) + relativeT) + T_Tr
        R = dot(R0, R_poseDiff)
        T = dot(R0, T_poseDiff) + T0
        out[i + 1, :, :3], out[i + 1, :, 3:] = R, T            # stored as float32 (np.array(poses, float32) at :272) ...
        R0, T0 = R, T                                             # ... but chained in the dtype the reference chains in
    return out.reshape(P + 1, 12)
