"""One caelo_nn_match call at 16384 x 16384 x 128 (BASELINE configs[3], SURVEY 8(d) distribution) between
cudaProfilerStart/Stop, for ncu:  ncu --profile-from-start off --set full --clock-control none -o R python tools/nn_ncu_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api
n, d = 16384, 128
rng = np.random.default_rng(n + d)
c0 = np.tanh(rng.standard_normal((n, d))).astype(np.float32)
perm = rng.permutation(n)
c1 = c0[perm].copy()
noisy = rng.random(n) < 0.4
c1[noisy] += (0.05 * rng.standard_normal((int(noisy.sum()), d))).astype(np.float32)
c1[~noisy] = np.tanh(rng.standard_normal((int((~noisy).sum()), d))).astype(np.float32)
ctx = api.default_context()
t0, t1 = torch.from_numpy(c0[None]).cuda(), torch.from_numpy(c1[None]).cuda()
for _ in range(3): idx = ctx.nn_match(t0, t1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
idx = ctx.nn_match(t0, t1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("known answers hit:", float((idx[0].cpu().numpy()[noisy] == perm[noisy]).mean()))
