#!/usr/bin/env python
"""bench.py — frame-pairs/s of the CAE-LO odometry hot path (keypts + desc + match + pose).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl ours|reference] [--no-extras]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

One step = one pass of the hot path over one batch of P consecutive synthetic frame pairs
(P+1 KITTI-seq-00-shaped frames: 64x1792x3 ring, 1024 keypoints, 3 x 16^3 voxel patches per
keypoint, 60-D descriptors, nn match + RANSAC + refit per pair) on every rank.  Frame pairs
shard across ranks with no data-path collective; ONE NCCL gather brings the per-pair poses of
all K steps to rank 0 (inside the timed region).  Prints ONE JSON line on rank 0.

Beside the headline the line carries (unless --no-extras): ``seq00`` — BASELINE configs[2], a whole
4541-frame synthetic drive through ``odometry.estimate_sequence``, the fixed 4540 pairs sharded over the N
ranks (strong scaling); ``nn_match`` — configs[3], the N x N x 128 argmin microbench; ``single_pair`` — the
latency of one pair (configs[1] unbatched); ``refine`` — configs[4]'s ICP refinement on the extended key points.

``--impl reference`` times the reference's own .py files (staged under baseline/_ref by
``__graft_entry__.build()``; Keras' predict restated in torch-CPU) on the host cores; the oracle port only if
the staged copy is missing.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frame-pairs/sec (keypts+desc+match+pose), KITTI-00 shape"
UNIT = "frame-pairs/s"
K_PTS = 1024
# algorithmic work per unit
FLOP_CONV12_PER_PATCH = 2 * (4096 * 27 * 8 + 512 * 216 * 16)                        # conv1+conv2: 5,308,416
FLOP_CONV3_PER_PATCH = 2 * 64 * 432 * 32                                            # 1,769,472
FLOP_DENSE_PER_PATCH = 2 * (2048 * 200 + 200 * 20)                                  # 827,200
BYTES_RESPOND_SELECT_PER_FRAME = 64 * 1792 * 3 * 4 + 69 * 1800 + 1024 * (12 + 16)   # fused: resp stays on chip
CPU_SAMPLE_PAIRS = 8                                                                # pairs per CPU pass (bounded sample)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: NVML polled every ~5 ms from a thread
    (nvidia_ml_py), falling back to `nvidia-smi -lms` if NVML is not importable."""

    def __init__(self, index: int):
        import threading
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._thr = None
        self._smi = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mx.append(float(mx))
                        r = int(get_reasons(h))
                        for k, bit in names.items():
                            if r & bit:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    self._stop.wait(0.005)

            self._thr = threading.Thread(target=poll, daemon=True)
            self._thr.start()
        except Exception:
            self._start_smi(index)

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def _start_smi(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self._smi = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self._smi = None

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join(timeout=2)
        elif self._smi is not None:
            self._smi.terminate()
            try:
                self._smi.wait(timeout=5)
            except Exception:
                self._smi.kill()
            self.f.flush()
            self.f.seek(0)
            for line in self.f.read().splitlines():
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    self.sm.append(float(c[1]))
                    self.mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            try:
                os.unlink(self.f.name)
            except OSError:
                pass
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (baseline/_ref, staged by build()) or, without it, the oracle port
# ------------------------------------------------------------------------------------------
def _cpu_init(kind, threads):
    """Pool worker start-up (spawned, like the reference's own workers: BatchPreprocess.py:237-238)."""
    import torch
    torch.set_num_threads(max(1, threads))
    if kind == "reference":
        from oracle import reference_stub
        reference_stub.load()                  # imports the unmodified reference modules (~6 s: Voxel.py:57-86)
        reference_stub.model("respond")
        reference_stub.model("encoder")
    else:
        from oracle import oracle
        oracle.build()


def _cpu_frame(args):
    kind, ring3, counter, v0, v1, v2 = args
    if kind == "reference":
        from oracle import reference_stub
        return reference_stub.frame_stage(ring3, counter, v0, v1, v2)
    from oracle import oracle
    resp = oracle.respond_predict(ring3[None])[0]
    kp, _px = oracle.select_keypoints(ring3, counter, resp)
    _, pl = oracle.get_patches_list(kp, v0, v1, v2)
    return kp, oracle.get_features_from_patches(pl)


def _cpu_pair(args):
    kind, pid, k0, c0, k1, c1 = args
    if kind == "reference":
        from oracle import reference_stub
        return reference_stub.pair_stage(pid, k0, c0, k1, c1)
    from oracle import oracle
    np.random.seed(pid)
    R, T, ok, i0, _i1, thr = oracle.solve_relative_pose(k0, c0, None, k1, c1, None)
    return np.r_[np.asarray(R, np.float32).ravel(), np.asarray(T, np.float32).ravel(), float(ok), len(i0), thr, 0]


class CpuArm:
    """The reference path on the host cores: a pool of worker processes (one frame / one pair per task, the way the
    reference fans frames out over processes: PoseEstimation.py:79-99, BatchPreprocess.py:194-228), torch threads
    inside each worker for ``predict``.  ``kind`` = "reference" when the staged / mounted reference tree is there
    (its own GetKeyPtsByAE / GetPatchesList / GetFeaturesFromPatches / SolveRelativePose run, numpy standing in for
    CuPy, torch-CPU for Keras), else "port" (the oracle restatement)."""

    def __init__(self, data):
        import multiprocessing as mp
        from oracle import reference_stub
        self.kind = "reference" if reference_stub.available() else "port"
        self.cores = os.cpu_count() or 1
        self.workers = max(1, min(self.cores, CPU_SAMPLE_PAIRS))
        self.threads = max(1, self.cores // self.workers)
        self.data = data
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_cpu_init, initargs=(self.kind, self.threads))
        self.frames = {}                      # frame index -> (KeyPts, Features)

    def close(self):
        self.pool.close()
        self.pool.join()

    def _job(self, f):
        d, off = self.data, self.data["vox_offsets"]
        v = [d["vox"][off[3 * f + s]:off[3 * f + s + 1]] for s in range(3)]
        return (self.kind, d["ring3"][f], d["counter"][f], *v)

    def run(self, frames, pairs):
        """Processes the NEW frames ``frames`` and then the pairs ``pairs`` = [(pair_id, f0, f1)] (their frames must
        be known by then) -> (seconds, pose rows [len(pairs),16])."""
        t0 = time.perf_counter()
        for f, r in zip(frames, self.pool.map(_cpu_frame, [self._job(f) for f in frames])):
            self.frames[f] = r
        rows = self.pool.map(_cpu_pair, [(self.kind, pid, *self.frames[a], *self.frames[b]) for pid, a, b in pairs])
        return time.perf_counter() - t0, np.asarray(rows, np.float32).reshape(len(pairs), 16)

    def describe(self, what):
        impl = ("the reference's own GetKeyPtsByAE / GetPatchesList / GetFeaturesFromPatches / SolveRelativePose "
                "(unmodified .py from baseline/_ref; numpy for CuPy, torch-CPU restatement of Keras predict)"
                if self.kind == "reference" else
                "oracle port of the reference path (C respond/select/match/RANSAC, scipy k-d tree patches, torch-CPU encoder)")
        return "%s; %s; %d worker processes x %d torch threads" % (what, impl, self.workers, self.threads)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  A step is a bounded sample
    of the P-pair step: CPU_SAMPLE_PAIRS new frames + the same number of pairs (steady state: every frame is processed
    once), walking back and forth over the step's P+1 frames."""
    if rank != 0:
        return
    from caelo_b200 import synth
    F = args.pairs + 1
    data = synth.make_frames(F, seed=1, device="cpu")
    arm = CpuArm(data)
    n = min(CPU_SAMPLE_PAIRS, args.pairs)
    walk = list(range(F)) + list(range(F - 2, 0, -1))                    # 0..P, P-1..1, 0..P, ...
    arm.run([walk[0]], [])                                                # the first frame of the drive (untimed)
    pos = 0
    times = []
    for step in range(args.warmup + args.steps):
        new = [walk[(pos + 1 + i) % len(walk)] for i in range(n)]
        prev = [walk[(pos + i) % len(walk)] for i in range(n)]
        sec, _rows = arm.run(new, [(min(a, b), a, b) for a, b in zip(prev, new)])
        pos += n
        if step >= args.warmup:
            times.append(sec)
    arm.close()
    total = sum(times)
    value = n * args.steps / total
    sample = arm.describe("%d of the step's %d pairs per step (%d new frames + %d pairs, every frame processed once)"
                          % (n, args.pairs, n, n))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _config(args, world):
    return {"workload": "configs[1] scaled to a batch: seq-00-shaped synthetic odometry, %d consecutive frame pairs "
                        "per step per GPU (%d frames of ~123k points from an HDL-64E-like beam pattern; 64x1792x3 ring, "
                        "1024 keypts/frame, 3x16^3 voxel patches, 60-D descriptors, 500-trial RANSAC)"
                        % (args.pairs, args.pairs + 1),
            "pairs_per_step_per_gpu": args.pairs, "keypoints": K_PTS, "parallelism": "pairs sharded x%d" % world,
            "l2": "inputs larger than L2: the steps cycle through 3 input sets (3 x 85 MB of ring images + voxel lists / "
                  "3 x 65 MB of raw scans vs 126 MB of L2); per-step intermediates (~1.7 GB of traffic) exceed L2 as well"}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def cpu_baseline_and_parity(data, gpu_rows, gpu_details):
    """N = 1 only: the CPU arm on a bounded sample of rank 0's step (pairs 0..n-1, every frame processed once), and —
    for free — an end-to-end parity check of the GPU rows against it:
      * key points of the sampled frames bit-identical,
      * descriptors inside the contract (|err| <= 1e-4 |ref| + 1e-5),
      * the GPU pose of every sampled pair bit-identical to the ORACLE's SolveRelativePose run on the GPU's own
        descriptors with the same seed (exactness of match + RANSAC + refit), and close to the CPU arm's pose (which
        starts from CPU descriptors, so an inlier or two may differ)."""
    from oracle import oracle
    arm = CpuArm(data)
    n = CPU_SAMPLE_PAIRS
    F = data["ring3"].shape[0]
    arm.run([F - 1 - i for i in range(n)] + [0], [])                     # warm-up pass on other frames (+ frame 0)
    sec, rows = arm.run(list(range(1, n + 1)), [(p, p, p + 1) for p in range(n)])
    arm.close()
    oracle.build()
    kp_g, ft_g = gpu_details["kpts"], gpu_details["feat"]
    # (a) against the CPU arm: its predict is a torch-CPU stand-in for Keras, NOT the arithmetic contract the kernels
    #     and the oracle share, so near-tied scores may order differently: report the overlap, compare descriptors on
    #     the key points both sides selected
    common, same_order, desc_err, desc_ok = [], True, 0.0, True
    for f in range(n + 1):
        kc, fc = arm.frames[f]
        idx = {tuple(r): i for i, r in enumerate(kc.view(np.uint32).tolist())}
        m = [(i, idx[tuple(r)]) for i, r in enumerate(kp_g[f].view(np.uint32).tolist()) if tuple(r) in idx]
        common.append(len(m) / max(1, kp_g[f].shape[0]))
        same_order &= kc.shape == kp_g[f].shape and bool(np.array_equal(kc, kp_g[f]))
        if m:
            gi, ci = np.array(m).T
            e = np.abs(ft_g[f][gi] - fc[ci])
            desc_err = max(desc_err, float(e.max()))
            desc_ok &= bool((e <= 1e-4 * np.abs(fc[ci]) + 1e-5).all())
    # (b) against the ORACLE (the contract arithmetic) on the first frames: key points bit-identical, descriptors
    #     inside the contract
    n_or = 3
    off = data["vox_offsets"]
    or_kp, or_desc = True, True
    for f in range(n_or):
        ko, _ = oracle.select_keypoints(data["ring3"][f], data["counter"][f], oracle.respond_predict(data["ring3"][f][None])[0])
        _, pl = oracle.get_patches_list(ko, *[data["vox"][off[3 * f + s]:off[3 * f + s + 1]] for s in range(3)])
        fo = oracle.get_features_from_patches(pl)
        or_kp &= bool(np.array_equal(ko, kp_g[f]))
        or_desc &= bool(ko.shape == kp_g[f].shape and (np.abs(ft_g[f] - fo) <= 1e-4 * np.abs(fo) + 1e-5).all())
    # (c) match + RANSAC + refit: the oracle's SolveRelativePose on the GPU's own key points + descriptors, same seed
    exact = 0
    for p in range(n):
        np.random.seed(p)
        R, T, ok, i0, _i1, thr = oracle.solve_relative_pose(kp_g[p], ft_g[p], None, kp_g[p + 1], ft_g[p + 1], None)
        g = gpu_rows[p]
        exact += int(np.array_equal(g[:9], np.asarray(R, np.float32).ravel()) and
                     np.array_equal(g[9:12], np.asarray(T, np.float32).ravel()) and
                     bool(g[12]) == bool(ok) and int(g[13]) == len(i0) and abs(g[14] - thr) < 1e-6)
    dpose = float(np.abs(gpu_rows[:n, :12] - rows[:, :12]).max())
    base = {"value": n / sec, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
            "sample": arm.describe("%d pairs (%d new frames) of rank 0's step, every frame processed once; 1 warm-up pass"
                                   % (n, n))}
    parity = {"parity_checked_pairs": n,
              "vs_oracle": {"frames": n_or, "keypoints_bit_exact": or_kp, "descriptors_within_contract": or_desc,
                            "pairs_pose_bit_exact_on_gpu_descriptors": exact},
              "vs_cpu_arm": {"keypoints_in_common_min": float(min(common)), "keypoints_identical_incl_order": bool(same_order),
                             "descriptors_within_contract_on_common_keypoints": bool(desc_ok),
                             "descriptor_max_abs_err": desc_err, "pose_max_abs_diff": dpose,
                             "inliers_gpu": [int(x) for x in gpu_rows[:n, 13]], "inliers_cpu_arm": [int(x) for x in rows[:, 13]]},
              "descriptor_contract": "|d - d_ref| <= 1e-4 |d_ref| + 1e-5 per component"}
    return base, parity


def nn_match_microbench(ctx, dev, peaks):
    """BASELINE configs[3]: N x N x D argmin, index-exact vs float64 cdist.  SURVEY §8(d) inputs: tanh(N(0,1)) rows;
    frame 1 = frame-0 rows permuted, 40 % with N(0, 0.05) noise (known answer), 60 % fresh draws."""
    import torch
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    out = []
    for n, d in [(1024, 60), (1024, 128), (2048, 128), (4096, 128), (8192, 128), (16384, 128)]:
        rng = np.random.default_rng(n + d)
        c0 = np.tanh(rng.standard_normal((n, d))).astype(np.float32)
        perm = rng.permutation(n)
        c1 = c0[perm].copy()
        noisy = rng.random(n) < 0.4
        c1[noisy] += (0.05 * rng.standard_normal((int(noisy.sum()), d))).astype(np.float32)
        c1[~noisy] = np.tanh(rng.standard_normal((int((~noisy).sum()), d))).astype(np.float32)
        t0, t1 = torch.from_numpy(c0[None]).to(dev), torch.from_numpy(c1[None]).to(dev)
        for _ in range(3):
            idx = ctx.nn_match(t0, t1)
        hit = float((idx[0].cpu().numpy()[noisy] == perm[noisy]).mean())
        ms = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.nn_match(t0, t1)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        t = float(np.median(ms)) * 1e-3
        flop, byts = 2.0 * n * n * d, (2 * n * d * 4 + n * 8)
        out.append({"shape": "%dx%dx%d" % (n, n, d), "ms": t * 1e3, "tflops_algorithmic": flop / t / 1e12,
                    "frac_of_bf16_sustained": flop / t / 1e12 / peaks["tf_sustained"],
                    "gbs_algorithmic": byts / t / 1e9, "frac_of_hbm": byts / t / 1e9 / peaks["hbm"],
                    "known_answers_hit": hit})
    return {"what": "configs[3]: descriptor NN match (cdist + argmin, index-exact), whole caelo_nn_match call "
                    "(operand prep + 3 split-fp16 tcgen05 passes + decide + exact re-scan), median of 10, L2 flushed",
            "shapes": out}


def single_pair_latency(pipe, data, dev):
    """configs[1] unbatched: ONE frame pair (2 frames) from ring images / voxel lists in HBM to the pose row on the
    host, per-call wall latency (the host waits for the result every call)."""
    import torch
    off = data["vox_offsets"]
    ring = torch.from_numpy(data["ring3"][:2]).to(dev)
    cnt = torch.from_numpy(data["counter"][:2]).to(dev)
    vox = torch.from_numpy(data["vox"][:off[6]]).to(dev)
    voff = off[:7].copy()
    for _ in range(5):
        pipe.run_device(ring, cnt, vox, voff, None, [0])
    lat = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.run_device(ring, cnt, vox, voff, None, [0])
        lat.append((time.perf_counter() - t0) * 1e3)
    return {"what": "one frame pair (2 frames resident in HBM) -> pose row on the host, wall clock per call",
            "ms_median": float(np.median(lat)), "ms_min": float(np.min(lat)), "pairs_per_s": 1e3 / float(np.median(lat))}


def seq00(pipe, args, rank, world, dev, dist):
    """BASELINE configs[2]: a whole seq-00-long synthetic drive (4541 frames, 4540 pairs) through
    ``odometry.estimate_sequence``: the FIXED pair list is sharded contiguously over the ranks (one-frame halo
    recomputed), 32-pair batches, ONE pose gather per sequence, pose chain on rank 0 — strong scaling."""
    import torch
    from caelo_b200 import odometry, pipeline, synth
    F = args.seq_frames
    P = F - 1
    lo, hi = pipeline.shard_pairs(P, rank, world)
    t0 = time.perf_counter()
    pts, off_local = synth.make_scans(hi - lo + 1, seed=7, first_frame=lo, device=dev)      # this rank's frames lo..hi
    gen_s = time.perf_counter() - t0

    def local_run(stacked):
        return odometry.estimate_sequence(stacked=stacked, stacked_first_frame=lo, n_frames=F, batch_pairs=32,
                                          rank=rank, world=world, pipe=pipe)

    def timed(stacked):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        poses, rel = local_run(stacked)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return float(sec.item()), poses, rel

    timed((pts, off_local))                                              # warm-up pass (scratch sizes, allocator)
    sec, poses, rel = timed((pts, off_local))
    out = None
    e2e = None
    # end to end: the same shard from pinned host memory (H2D of batch i+1 under batch i's kernels)
    nbytes = pts.numel() * 4
    try:
        import psutil
        room = psutil.virtual_memory().available
    except Exception:
        room = 0
    if room > 3 * nbytes * max(1, world):
        host = torch.empty(pts.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(pts)
        torch.cuda.synchronize()
        sec_h, poses_h, rel_h = timed((host, off_local))
        e2e = {"value": P / sec_h, "unit": UNIT, "seconds": sec_h, "h2d_bytes": int(nbytes) if world == 1 else None,
               "same_poses_as_device_resident": bool(rank != 0 or np.array_equal(rel_h, rel))}
    if rank == 0:
        # trajectory against the known motion of the synthetic drive
        gt = []
        for f in range(F):
            R, T = synth.sensor_pose(f)
            gt.append(np.c_[R, T.reshape(3, 1)].reshape(12))
        gt = np.asarray(gt, np.float64)
        rre, rte = [], []
        for i in range(P):
            R, T = odometry.GetRelRtBetween2Poses(poses[i].astype(np.float64), poses[i + 1].astype(np.float64))
            Rg, Tg = odometry.GetRelRtBetween2Poses(gt[i], gt[i + 1])
            c = np.clip((np.trace(np.dot(Rg.T, R)) - 1) / 2, -1, 1)
            rre.append(np.degrees(np.arccos(c)))
            rte.append(np.linalg.norm(T - Tg))
        rre, rte = np.asarray(rre), np.asarray(rte)
        succ = float(((rre < 1.0) & (rte < 0.5)).mean())             # EvaluationOnRegistration.py:23-24
        drift = float(np.linalg.norm(poses[-1].reshape(3, 4)[:, 3] - gt[-1].reshape(3, 4)[:, 3]))
        out = {"what": "configs[2]: %d-frame synthetic drive (seq 00 has 4541), %d pairs sharded contiguously over %d "
                       "rank(s), 32-pair batches from raw scans resident in HBM, one pose gather, pose chain on rank 0"
                       % (F, P, world),
               "scaling": "strong", "pairs": P, "seconds": sec, "value": P / sec, "unit": UNIT,
               "e2e": e2e, "pairs_with_model": int((rel[:, 12] != 0).sum()),
               "rre_deg_mean": float(rre.mean()), "rte_m_mean": float(rte.mean()), "registration_success": succ,
               "end_point_drift_m": drift, "path_length_m": float(0.7 * P),
               "scan_generation_s": gen_s}
    return out


def refine_bench(ctx, pipe, data, dev, rank, world, dist):
    """configs[4]: RefinePoses.py's ICP refinement on the extended key points, all pairs of the step in one batched
    device-side ICP (caelo_b200.refine)."""
    try:
        from caelo_b200 import refine
    except ImportError:
        return None
    return refine.bench(ctx, pipe, data, dev, rank, world, dist)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: caelo_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    from caelo_b200 import api, pipeline, synth

    ctx = api.Context(local_rank)
    pipe = pipeline.OdometryPipeline(ctx, K_PTS)
    P = args.pairs
    F = P + 1
    K = args.steps
    # pair ids of this rank inside a notional sequence: rank r owns pairs [r*P, (r+1)*P)
    pair_ids = list(range(rank * P, (rank + 1) * P))
    # INPUTS LARGER THAN L2: the steps cycle through N_SETS different 33-frame stretches of the drive (3 x 85 MB of ring
    # images + voxel lists, 3 x 65 MB of raw scans, against 126 MB of L2); a set comes round again after the two others and
    # ~5 GB of intermediate traffic.  Set 0 is the one the parity check, the CPU baseline and the sub-runs use.
    N_SETS = 3

    def make_set(s):
        d = synth.make_frames(F, seed=1 + rank + 1000 * s, first_frame=(rank + s * world) * P, device=dev)
        h = dict(ring=torch.from_numpy(d["ring3"]).pin_memory(), counter=torch.from_numpy(d["counter"]).pin_memory(),
                 vox=torch.from_numpy(d["vox"]).pin_memory())
        so = np.zeros(F + 1, np.int64)
        so[1:] = np.cumsum([x.shape[0] for x in d["scans"]])
        sh = torch.from_numpy(np.concatenate(d["scans"], 0)).pin_memory()
        return dict(data=d, host=h, voff=d["vox_offsets"], dev=tuple(h[k].to(dev) for k in ("ring", "counter", "vox")),
                    soff=so, scans_h=sh, d_scans=sh.to(dev),
                    pair_ids=list(range((rank + s * world) * P, (rank + s * world + 1) * P)))

    sets = [make_set(s) for s in range(N_SETS)]
    data, host, voff, soff, scans_h, pair_ids = (sets[0][k] for k in ("data", "host", "voff", "soff", "scans_h", "pair_ids"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def enqueue_rings(k=0):
        z = sets[k % N_SETS]
        return pipe.enqueue_device(*z["dev"], z["voff"], None, z["pair_ids"])

    def enqueue_scans(k=0):
        z = sets[k % N_SETS]
        return pipe.enqueue_device_scans(z["d_scans"], z["soff"], None, z["pair_ids"])

    def timed_steps(enqueue):
        """K steps queued back to back (nothing waits for the device in between; the pairs stage of step i runs on the
        pipeline's tail stream next to the frame stages of step i+1), step k on input set k mod 3 (inputs larger than L2),
        ONE pair of CUDA events around all K steps; then the host reads every step's result rows and ONE gather brings
        all K x P rows to rank 0.  -> (device time of the K steps + the gather, wall seconds incl. collect + gather, rows
        on this rank, gathered rows on rank 0)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        handles = []
        ev0.record()
        for k in range(K):
            handles.append(enqueue(k))
        pipe.join()                  # the pairs stages run on the pipeline's tail stream: the end event covers them too
        ev1.record()
        rows = [pipe.collect(h) for h in handles]
        tg0 = time.perf_counter()
        allrows = pipeline.gather_poses(np.concatenate(rows, 0), dev, cap=K * P)
        gather_s = time.perf_counter() - tg0
        barrier()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1) + gather_s * 1e3
        return ms, wall, rows, allrows

    def reduce_max(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    # ---- warm-up ----
    pipe.keep_details = True
    for _ in range(max(args.warmup, 3)):
        poses = pipe.collect(enqueue_rings())
    det = pipe.last_details
    gpu_details = {"kpts": det["kpts"].cpu().numpy(), "feat": det["feat"].cpu().numpy()} if rank == 0 else None
    gpu_rows0 = poses
    pipe.keep_details = False
    pipe.last_details = None
    ok_pairs = int((poses[:, 12] != 0).sum())
    for _ in range(2):
        poses_s = pipe.collect(enqueue_scans())
    same = bool(np.array_equal(poses_s, poses))
    pipeline.gather_poses(np.concatenate([poses] * K, 0), dev, cap=K * P)      # warm-up of the collective (NCCL sets its communicator up lazily)

    # ---- timed region: K steps from ring images + voxel lists resident in HBM ----
    # one untimed round in exactly the shape of the timed one (K steps queued ahead, handles alive): afterwards the caching
    # allocator owns every block the timed round needs — a cudaMalloc inside the timed region synchronises the device
    timed_steps(enqueue_rings)
    timed_steps(enqueue_scans)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.profile(True)
    ctx.profile_fetch()
    launches0 = ctx.launches
    ms, _wall, _rows, _all = timed_steps(enqueue_rings)
    launches = ctx.launches - launches0
    prof = ctx.profile_fetch()
    ctx.profile(False)
    # ---- the same from raw scans resident in HBM (f1 / f2+a6 on the device in front of the hot path) ----
    ctx.profile(True)
    ctx.profile_fetch()
    ms_s, _w, _r, _a = timed_steps(enqueue_scans)
    prof_s = ctx.profile_fetch()
    ctx.profile(False)
    ms_total, ms_scans = reduce_max(ms, ms_s)

    # ---- end to end: host (pinned) inputs, H2D + kernels + D2H of every step's rows, one gather at the end ----
    def e2e_run(kind):
        def steps(n):
            for k in range(n):
                z = sets[k % N_SETS]
                yield (("rings", z["host"]["ring"], z["host"]["counter"], z["host"]["vox"], z["voff"], z["pair_ids"])
                       if kind == "rings" else ("scans", z["scans_h"], z["soff"], z["pair_ids"]))
        for _p in pipe.run_host_stream(steps(3)):
            pass
        barrier()
        t0 = time.perf_counter()
        rows = list(pipe.run_host_stream(steps(K)))
        pipeline.gather_poses(np.concatenate(rows, 0), dev, cap=K * P)
        barrier()
        return time.perf_counter() - t0

    e2e_rings_s, e2e_scans_s = reduce_max(e2e_run("rings"), e2e_run("scans"))
    clocks = sampler.stop() if sampler else None
    # bytes per step, averaged over the steps actually run (the input sets differ a little in voxel / point counts)
    used = [sets[k % N_SETS] for k in range(K)]
    h2d_rings = sum(sum(z["host"][k].numel() * z["host"][k].element_size() for k in z["host"]) + z["voff"].nbytes + 8 * P
                    for z in used) / K
    h2d_scans = sum(z["scans_h"].numel() * 4 + z["soff"].nbytes + 8 * P for z in used) / K
    d2h = P * 32 * 4

    extras = {}
    if not args.no_extras:
        def extra(name, fn, *a):
            # a sub-run never takes the headline down (every rank runs the same sub-runs, so collectives still pair up)
            try:
                extras[name] = fn(*a)
            except Exception as e:                                      # noqa: BLE001
                extras[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        extra("seq00", seq00, pipe, args, rank, world, dev, dist)
        if rank == 0:
            extra("single_pair", single_pair_latency, pipe, data, dev)
            extra("nn_match", nn_match_microbench, ctx, dev, _peaks())
        extra("refine", refine_bench, ctx, pipe, data, dev, rank, world, dist)

    if rank == 0:
        peaks = _peaks()
        n_patches = F * 3 * K_PTS

        def kern(name, work, peak, unit, scale):
            if name not in prof:
                return None
            n, tot = prof[name]
            sec = tot / n * 1e-3
            ach = work / sec / scale
            return {"kernel": name, "launches_per_step": n // K, "avg_ms": tot / n,
                    "share_of_step": tot / ms if ms else None, "achieved": ach, "peak": peak, "unit": unit,
                    "frac": ach / peak}

        kernels = [k for k in (
            kern("conv12_pair_kernel", FLOP_CONV12_PER_PATCH * n_patches, peaks["tf_sustained"], "TFLOP/s", 1e12),
            kern("conv12_tc_kernel", FLOP_CONV12_PER_PATCH * n_patches, peaks["tf_sustained"], "TFLOP/s", 1e12),
            kern("conv3_oct_kernel", FLOP_CONV3_PER_PATCH * n_patches, peaks["tf_sustained"], "TFLOP/s", 1e12),
            kern("conv3_tc_kernel", FLOP_CONV3_PER_PATCH * n_patches, peaks["tf_sustained"], "TFLOP/s", 1e12),
            kern("dense_tc_kernel", FLOP_DENSE_PER_PATCH * n_patches, peaks["tf_sustained"], "TFLOP/s", 1e12),
            kern("respond_score_kernel<fused>", BYTES_RESPOND_SELECT_PER_FRAME * F, peaks["hbm"], "GB/s", 1e9),
            kern("gather_kernel", (n_patches * 512 + F * 3 * 4), peaks["hbm"], "GB/s", 1e9),
            kern("nn_tc_kernel", 2.0 * P * K_PTS * K_PTS * 60, peaks["tf_sustained"], "TFLOP/s", 1e12),
        ) if k]
        top = max(prof.items(), key=lambda kv: kv[1][1])[0] if prof else None
        dom = next((k for k in kernels if k["kernel"] == top), kernels[0] if kernels else None)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if dom and os.path.isfile(tp):
            traffic = json.load(open(tp)).get(dom["kernel"])
        roofline = None
        if dom:
            roofline = {"bound": "tensor" if dom["unit"] == "TFLOP/s" else "hbm", "achieved": dom["achieved"],
                        "peak": dom["peak"], "unit": dom["unit"], "frac": dom["frac"], "traffic": traffic,
                        "kernel": dom["kernel"], "avg_launch_ms": dom["avg_ms"], "share_of_step": dom["share_of_step"],
                        "peak_source": peaks["source"] + (" bf16 sustained" if dom["unit"] == "TFLOP/s" else " copy"),
                        "frac_of_burst_peak": dom["achieved"] / peaks["tf_burst"] if dom["unit"] == "TFLOP/s" else None,
                        "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of the "
                                          "round's ncu --set full capture of this kernel, per launch)",
                        "note": "per-kernel times are CUDA-event intervals on the launching stream; the pairs stage (draw_samples "
                                "... kabsch) runs on the pipeline's tail stream NEXT TO the following step's frame stages, so its "
                                "intervals include waiting for SMs and the per-kernel times add up to more than the step",
                        "all_kernels": kernels,
                        "time_by_kernel_ms_per_step": {k: v[1] / K for k, v in prof.items()}}
        e2e_best = min(e2e_rings_s, e2e_scans_s)
        from_rings = e2e_rings_s <= e2e_scans_s
        line = {"metric": METRIC, "value": world * P * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": _config(args, world), "clocks": clocks,
                "e2e": {"value": world * P * K / e2e_best, "unit": UNIT,
                        "h2d_bytes_per_step": int(h2d_rings if from_rings else h2d_scans), "d2h_bytes_per_step": int(d2h),
                        "input": "ring images + voxel lists" if from_rings else "raw scans",
                        "how": "pipeline.run_host_stream over the K steps (pinned host inputs; batch i+1's H2D overlaps "
                               "batch i's kernels; every step's result rows are read back), one pose gather at the end",
                        "from_ring_images": {"value": world * P * K / e2e_rings_s, "h2d_bytes_per_step": int(h2d_rings)},
                        "from_raw_scans": {"value": world * P * K / e2e_scans_s, "h2d_bytes_per_step": int(h2d_scans)}},
                "gpu_launches": int(launches), "roofline": roofline,
                "pairs_with_model": ok_pairs,
                "from_scans": {"what": "same step started from the raw (N,4) scans resident in HBM: ProjectPC2SphericalRing + "
                                       "Voxelization's voxel sets computed on the device (f1, f2+a6 fused) in front of the hot path",
                               "value": world * P * K / (ms_scans * 1e-3), "unit": UNIT, "ms_per_step": ms_scans / K,
                               "poses_identical_to_ring_path": same,
                               "time_by_kernel_ms_per_step": {k: v[1] / K for k, v in prof_s.items()}}}
        line.update({k: v for k, v in extras.items() if v is not None})
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], line["parity"] = cpu_baseline_and_parity(data, gpu_rows0, gpu_details)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=32, help="frame pairs per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the seq00 / nn_match / single_pair / refine sub-runs")
    ap.add_argument("--seq-frames", type=int, default=4541, help="frames of the configs[2] drive (seq 00 has 4541)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
