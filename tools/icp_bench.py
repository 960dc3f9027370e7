"""f4 measurement: api.ICP (device 1-NN + Kabsch + update, host loop control) vs the CPU oracle restatement of
MyICP.ICP on extended-key-point-sized synthetic sets.  One JSON line per size.
    python tools/icp_bench.py [--out profiles/r1_icp_bench.jsonl]"""
import argparse, contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api
from oracle import oracle

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); args = ap.parse_args()
ctx = api.default_context()
lines = []
for n in (12000, 50000):
    rng = np.random.default_rng(n)
    # points on a few planes (ground + walls), frame 1 = frame 0 moved by a small rigid motion + noise
    g = np.c_[rng.uniform(-40, 40, (n // 2, 2)), np.full(n // 2, -1.7)]
    w = np.c_[rng.uniform(-40, 40, n - n // 2), np.full(n - n // 2, 12.0), rng.uniform(-1.7, 4, n - n // 2)]
    p0 = np.r_[g, w].astype(np.float32)
    a = np.deg2rad(0.4)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    p1 = ((p0[rng.permutation(n)] - [0.12, 0.05, 0.0]) @ R + rng.normal(0, 0.01, (n, 3))).astype(np.float32)
    kw = dict(inlierThreshold=0.3, smallShiftThreshold=0.1, ep=0.01)
    gi, oi = {}, {}
    with contextlib.redirect_stdout(io.StringIO()):
        api.ICP(p0, p1, info=gi, **kw)                    # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        Rg, Tg, okg = api.ICP(p0, p1, info=gi, **kw)
        torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
    ctx.profile(True); ctx.profile_fetch()
    with contextlib.redirect_stdout(io.StringIO()):
        api.ICP(p0, p1, **kw)
    prof = ctx.profile_fetch(); ctx.profile(False)
    t0 = time.perf_counter(); Ro, To, oko = oracle.icp(p0, p1, info=oi, **kw); t_cpu = time.perf_counter() - t0
    nn_calls, nn_ms = prof["nn3_kernel"]
    line = {"workload": "ICP %d x %d points (MyICP.ICP arguments of RefinementCore-like use)" % (n, n), "iters": gi["iters"],
            "gpu_ms": t_gpu * 1e3, "gpu_ms_per_iter": t_gpu * 1e3 / gi["iters"], "nn3_kernel_ms_per_iter": nn_ms / nn_calls,
            "nn3_f64_gflops": 8.0 * n * n / (nn_ms / nn_calls * 1e-3) / 1e9,
            "cpu_oracle_ms": t_cpu * 1e3, "speedup": t_cpu / t_gpu,
            "identical_to_oracle": bool(okg == oko and np.array_equal(Rg, Ro) and np.array_equal(Tg, To) and gi == oi)}
    print(json.dumps(line)); lines.append(line)
if args.out:
    open(args.out, "w").write("\n".join(json.dumps(l) for l in lines) + "\n")
