"""BASELINE configs[3]: descriptor NN-match microbench, N x N x 128 (and x 60) argmin on one GPU.
Prints one JSON line per shape: device time of caelo_nn_match (CUDA events, L2 flushed between runs), the
algorithmic 2*N*M*D FLOP rate against the measured bf16 tensor peak and the algorithmic bytes against HBM.
    python tools/nn_microbench.py [--out profiles/r1_nn_microbench.jsonl]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); args = ap.parse_args()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.isfile("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
ctx = api.default_context()
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
lines = []
for n, d in [(1024, 60), (1024, 128), (2048, 128), (4096, 128), (8192, 128), (16384, 128)]:
    rng = np.random.default_rng(n)
    c0 = np.tanh(rng.standard_normal((n, d))).astype(np.float32)
    c1 = c0[rng.permutation(n)] + (0.05 * rng.standard_normal((n, d))).astype(np.float32)
    t0, t1 = torch.from_numpy(c0[None]).cuda(), torch.from_numpy(c1[None]).cuda()
    for _ in range(3): ctx.nn_match(t0, t1)
    ms = []
    for _ in range(10):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctx.nn_match(t0, t1); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    t = float(np.median(ms)) * 1e-3
    flop, byts = 2.0 * n * n * d, (2 * n * d * 4 + n * 8)
    line = {"workload": "nn-match %dx%dx%d (cdist+argmin, index-exact)" % (n, n, d), "ms": t * 1e3,
            "tflops_algorithmic": flop / t / 1e12, "frac_of_bf16_sustained": flop / t / 1e12 / peaks["bf16_tflops_sustained"],
            "gbs_algorithmic": byts / t / 1e9, "frac_of_hbm": byts / t / 1e9 / peaks["hbm_gbs"],
            "note": "three split-fp16 MMA passes + norm / decide / exact re-scan kernels inside the timed call"}
    print(json.dumps(line)); lines.append(line)
if args.out:
    open(args.out, "w").write("\n".join(json.dumps(l) for l in lines) + "\n")
