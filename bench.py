#!/usr/bin/env python
"""bench.py — frame-pairs/s of the CAE-LO odometry hot path (keypts + desc + match + pose).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

One step = one pass of the hot path over one batch of P consecutive synthetic frame pairs
(P+1 KITTI-seq-00-shaped frames: 64x1792x3 ring, 1024 keypoints, 3 x 16^3 voxel patches per
keypoint, 60-D descriptors, nn match + RANSAC + refit per pair) on every rank.  Frame pairs
shard across ranks with no data-path collective; one NCCL gather brings the per-pair poses to
rank 0 (inside the timed region).  Prints ONE JSON line on rank 0.

``--impl reference`` times the CPU restatement of the reference path (oracle/, the reference's
own .py cannot travel to the GPU box and Keras/TF/CuPy are not installable) on the host cores.
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
# CPU arm (oracle port of the reference path)
# ------------------------------------------------------------------------------------------
def _cpu_frame(args):
    ring3, counter, v0, v1, v2 = args
    from oracle import oracle
    resp = oracle.respond_predict(ring3[None])[0]
    kp, _px = oracle.select_keypoints(ring3, counter, resp)
    _, pl = oracle.get_patches_list(kp, v0, v1, v2)
    return kp, pl


def _cpu_pair(args):
    pid, k0, c0, k1, c1 = args
    from oracle import oracle
    np.random.seed(pid)
    R, T, ok, i0, i1, thr = oracle.solve_relative_pose(k0, c0, None, k1, c1, None)
    return np.r_[np.asarray(R, np.float32).ravel(), np.asarray(T, np.float32).ravel(), float(ok), len(i0), thr, 0]


def cpu_pass(data, n_frames, pool, threads):
    """Steady-state CPU pass over n_frames frames / n_frames-1 pairs; returns seconds."""
    import torch
    from oracle import oracle
    torch.set_num_threads(threads)
    off = data["vox_offsets"]
    jobs = []
    for f in range(n_frames):
        v = [data["vox"][off[3 * f + s]:off[3 * f + s + 1]] for s in range(3)]
        jobs.append((data["ring3"][f], data["counter"][f], *v))
    t0 = time.perf_counter()
    frames = pool.map(_cpu_frame, jobs) if pool else [_cpu_frame(j) for j in jobs]
    feats = [oracle.get_features_from_patches(pl) for _, pl in frames]   # torch-CPU, all threads (Keras stand-in)
    pj = [(p, frames[p][0], feats[p], frames[p + 1][0], feats[p + 1]) for p in range(n_frames - 1)]
    _poses = pool.map(_cpu_pair, pj) if pool else [_cpu_pair(j) for j in pj]
    return time.perf_counter() - t0


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation (oracle port) on the host cores."""
    if rank != 0:
        return
    import multiprocessing as mp
    from caelo_b200 import synth
    from oracle import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    n_frames = min(args.pairs, 4) + 1            # bounded sample: 4 pairs of the P-pair step
    data = synth.make_frames(n_frames, seed=0)
    workers = min(cores, n_frames)
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for _ in range(args.warmup):
            cpu_pass(data, n_frames, pool, cores)
        t = [cpu_pass(data, n_frames, pool, cores) for _ in range(args.steps)]
    total = sum(t)
    value = (n_frames - 1) * args.steps / total
    sample = ("%d of the step's %d pairs (%d synthetic frames) per step; oracle port of the reference path "
              "(C respond/select/match/RANSAC, scipy k-d tree patches, torch-CPU encoder), %d worker processes "
              "+ %d torch threads" % (n_frames - 1, args.pairs, n_frames, workers, cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _config(args, world):
    return {"workload": "configs[1] scaled to a batch: seq-00-shaped synthetic odometry, %d consecutive frame pairs "
                        "per step per GPU (%d frames; 64x1792x3 ring, 1024 keypts/frame, 3x16^3 voxel patches, "
                        "60-D descriptors, 500-trial RANSAC)" % (args.pairs, args.pairs + 1),
            "pairs_per_step_per_gpu": args.pairs, "keypoints": K_PTS, "parallelism": "pairs sharded x%d" % world,
            "l2": "256 MiB write between timed steps (untimed); per-step intermediates (~0.8 GB) exceed L2"}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
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
    # pair ids of this rank inside a notional sequence: rank r owns pairs [r*P, (r+1)*P)
    pair_ids = list(range(rank * P, (rank + 1) * P))
    data = synth.make_frames(F, seed=1 + rank, first_frame=rank * P)
    host = dict(ring=torch.from_numpy(data["ring3"]).pin_memory(),
                counter=torch.from_numpy(data["counter"]).pin_memory(),
                vox=torch.from_numpy(data["vox"]).pin_memory())
    voff = data["vox_offsets"]
    d_ring, d_counter, d_vox = (host[k].to(dev) for k in ("ring", "counter", "vox"))
    d_samples = None      # RANSAC sample indices are generated on the device inside every step (ctx.draw_samples)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        poses = pipe.run_device(d_ring, d_counter, d_vox, voff, d_samples, pair_ids)
        return pipeline.gather_poses(poses, dev, cap=P)

    def step_host():
        poses = pipe.run_host(host["ring"], host["counter"], host["vox"], voff, pair_ids)
        return pipeline.gather_poses(poses, dev, cap=P)

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        poses = step_device()
    ok_pairs = None
    if rank == 0 and poses is not None:
        ok_pairs = int((poses[:, 12] != 0).sum())

    # ---- timed region: K steps, CUDA events on the launch stream, L2 flushed between steps ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.profile(True)
    ctx.profile_fetch()
    launches0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()
        torch.cuda.synchronize()
        a.record()
        step_device()
        b.record()
    barrier()
    launches = ctx.launches - launches0
    prof = ctx.profile_fetch()
    ctx.profile(False)
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())

    # ---- end to end: host (pinned) inputs, H2D + sample drawing + kernels + D2H every step ----
    # (the public sequence call: pipeline.run_host_stream uploads batch i+1 while batch i computes; every step
    #  still copies its own inputs from pinned host memory and reads its own result back)
    def host_steps(n):
        for _ in range(n):
            yield ("rings", host["ring"], host["counter"], host["vox"], voff, pair_ids)

    for poses in pipe.run_host_stream(host_steps(3)):
        pipeline.gather_poses(poses, dev, cap=P)
    barrier()
    t0 = time.perf_counter()
    for poses in pipe.run_host_stream(host_steps(args.steps)):
        pipeline.gather_poses(poses, dev, cap=P)
    barrier()
    e2e_s = time.perf_counter() - t0
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        step_host()
    barrier()
    e2e_single_s = (time.perf_counter() - t0) / 5
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if sampler else None

    # ---- same job started one stage earlier: raw scans -> poses ----
    from_scans = None
    if not args.no_from_scans:
        soff = np.zeros(F + 1, np.int64)
        soff[1:] = np.cumsum([s.shape[0] for s in data["scans"]])
        scans_h = torch.from_numpy(np.concatenate(data["scans"], 0)).pin_memory()
        d_scans = scans_h.to(dev)
        for _ in range(3):
            poses_s = pipe.run_device_scans(d_scans, soff, d_samples, pair_ids)
        same = bool(np.array_equal(poses_s, pipe.run_device(d_ring, d_counter, d_vox, voff, d_samples, pair_ids)))
        ctx.profile(True)
        ctx.profile_fetch()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for a, b in evs:
            flush.zero_()
            torch.cuda.synchronize()
            a.record()
            pipeline.gather_poses(pipe.run_device_scans(d_scans, soff, d_samples, pair_ids), dev, cap=P)
            b.record()
        barrier()
        prof_s = ctx.profile_fetch()
        ctx.profile(False)
        ms_s = sum(a.elapsed_time(b) for a, b in evs)
        def scan_steps(n):
            for _ in range(n):
                yield ("scans", scans_h, soff, pair_ids)

        for poses in pipe.run_host_stream(scan_steps(3)):
            pipeline.gather_poses(poses, dev, cap=P)
        barrier()
        t0 = time.perf_counter()
        for poses in pipe.run_host_stream(scan_steps(args.steps)):
            pipeline.gather_poses(poses, dev, cap=P)
        barrier()
        e2e_scans_s = time.perf_counter() - t0
        t = torch.tensor([ms_s, e2e_scans_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        from_scans = {"what": "same step started from the raw (N,4) scans: ProjectPC2SphericalRing + Voxelization's "
                              "voxel sets computed on the device (f1, f2+a6 fused) in front of the hot path",
                      "value": world * P * args.steps / (float(t[0].item()) * 1e-3), "unit": UNIT,
                      "ms_per_step": float(t[0].item()) / args.steps,
                      "e2e": {"value": world * P * args.steps / float(t[1].item()), "unit": UNIT,
                              "h2d_bytes_per_step": int(scans_h.numel() * 4 + soff.nbytes + 8 * P), "d2h_bytes_per_step": P * 32 * 4},
                      "poses_identical_to_ring_path": same,
                      "time_by_kernel_ms_per_step": {k: v[1] / args.steps for k, v in prof_s.items()}}

    if rank == 0:
        peaks = _peaks()
        n_patches = F * 3 * K_PTS

        def kern(name, work, peak, unit, scale):
            if name not in prof:
                return None
            n, tot = prof[name]
            sec = tot / n * 1e-3
            ach = work / sec / scale
            return {"kernel": name, "launches_per_step": n // args.steps, "avg_ms": tot / n,
                    "share_of_step": tot / ms if ms else None, "achieved": ach, "peak": peak, "unit": unit,
                    "frac": ach / peak}

        kernels = [k for k in (
            kern("conv12_tc_kernel", FLOP_CONV12_PER_PATCH * n_patches, peaks["tf_sustained"], "TFLOP/s", 1e12),
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
                        "all_kernels": kernels,
                        "time_by_kernel_ms_per_step": {k: v[1] / args.steps for k, v in prof.items()}}
        h2d = sum(host[k].numel() * host[k].element_size() for k in host) + voff.nbytes + 8 * P
        d2h = P * 32 * 4
        line = {"metric": METRIC, "value": world * P * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": _config(args, world), "clocks": clocks,
                "e2e": {"value": world * P * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h),
                        "how": "pipeline.run_host_stream over the K steps (pinned host inputs; batch i+1's H2D overlaps "
                               "batch i's kernels); one isolated run_host call (no overlap between calls) gives "
                               "%.1f %s" % (world * P / e2e_single_s, UNIT)},
                "gpu_launches": int(launches), "roofline": roofline,
                "pairs_with_model": ok_pairs, "from_scans": from_scans}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(data)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(data):
    """Oracle port of the reference path on the host cores, bounded sample (3 frames / 2 pairs)."""
    import multiprocessing as mp
    from oracle import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    n_frames = 3
    with mp.get_context("fork").Pool(min(cores, n_frames)) as pool:
        cpu_pass(data, n_frames, pool, cores)
        sec = cpu_pass(data, n_frames, pool, cores)
    return {"value": (n_frames - 1) / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "2 pairs (3 frames) of the same synthetic workload, steady-state accounting "
                      "(each frame processed once); oracle port, %d worker processes + %d torch threads; "
                      "1 warm-up pass" % (min(cores, n_frames), cores)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=32, help="frame pairs per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-from-scans", action="store_true")
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
