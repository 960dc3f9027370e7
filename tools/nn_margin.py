"""How accurate is the tensor-core pass of the nn match?  Measures |d2_pass - d2_exact| of every column's best row in units of
|a|max |b_j| (the quantity the decision margin E is a multiple of) on several descriptor distributions, and the share of
columns that go to the exact re-scan for a range of margins:  python tools/nn_margin.py"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api

ctx = api.default_context()
def dataset(kind, n, d, rng):
    if kind == "tanh":      # 40 % noisy copies, 60 % fresh
        c0 = np.tanh(rng.standard_normal((n, d))).astype(np.float32)
        c1 = c0[rng.permutation(n)].copy(); noisy = rng.random(n) < 0.4
        c1[noisy] += (0.05 * rng.standard_normal((int(noisy.sum()), d))).astype(np.float32)
        c1[~noisy] = np.tanh(rng.standard_normal((int((~noisy).sum()), d))).astype(np.float32)
    elif kind == "positive":  # all components positive: the accumulator only grows (worst case for truncation)
        c0 = rng.random((n, d)).astype(np.float32); c1 = rng.random((n, d)).astype(np.float32)
    elif kind == "big":       # norms far from 1
        c0 = (rng.standard_normal((n, d)) * 37).astype(np.float32); c1 = (rng.standard_normal((n, d)) * 37).astype(np.float32)
    else:                      # tiny values: fp16 lo parts are subnormal
        c0 = (rng.standard_normal((n, d)) * 1e-3).astype(np.float32); c1 = (rng.standard_normal((n, d)) * 1e-3).astype(np.float32)
    return c0, c1
rng = np.random.default_rng(0)
worst = 0.0
for kind in ("tanh", "positive", "big", "tiny"):
    for n, d in ((1024, 60), (4096, 128), (16384, 128), (2048, 33)):
        c0, c1 = dataset(kind, n, d, rng)
        t0, t1 = torch.from_numpy(c0[None]).cuda(), torch.from_numpy(c1[None]).cuda()
        line = {"data": kind, "shape": "%dx%dx%d" % (n, n, d)}
        for lg in (13, 15, 16, 17):
            ctx.check(ctx.lib.caelo_debug_set_nn_margin(ctx.h, ctypes.c_float(2.0 ** -lg)))
            idx = ctx.nn_match(t0, t1)
            bd = torch.empty((1, n), dtype=torch.float32, device="cuda"); bi = torch.empty((1, n), dtype=torch.int32, device="cuda")
            nu = torch.zeros((1,), dtype=torch.int32, device="cuda")
            ctx.check(ctx.lib.caelo_debug_nn_last(ctx.h, 1, n, n, api._ptr(bd), None, api._ptr(bi), api._ptr(nu), api._stream()))
            line["undecided@2^-%d" % lg] = int(nu.item()) / n
            if lg == 13:
                ref_idx = idx.clone()
                b = bi[0].cpu().numpy().astype(np.int64)
                diff = c0[b].astype(np.float64) - c1.astype(np.float64)
                exact = (diff * diff).sum(1)
                scale = np.sqrt((c0.astype(np.float64) ** 2).sum(1).max() * (c1.astype(np.float64) ** 2).sum(1))
                err = np.abs(bd[0].cpu().numpy().astype(np.float64) - exact) / scale
                line["max_err_log2"] = float(np.log2(err.max() + 1e-300)); worst = max(worst, err.max())
            else:
                line["same_indices@2^-%d" % lg] = bool(torch.equal(idx, ref_idx))
        print(json.dumps(line), flush=True)
print(json.dumps({"worst_err_log2": float(np.log2(worst))}))
ctx.check(ctx.lib.caelo_debug_set_nn_margin(ctx.h, ctypes.c_float(2.0 ** -13)))
