"""Where one batched step spends its time on the stream (CUDA events between stages) and on the host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api, pipeline, synth

P = 32
ctx = api.default_context(); pipe = pipeline.OdometryPipeline(ctx)
d = synth.make_frames(P + 1, seed=1)
ring, cnt, vox = (torch.from_numpy(d[k]).cuda() for k in ("ring3", "counter", "vox"))
voff = d["vox_offsets"]; pair_ids = list(range(P))
smp = torch.from_numpy(pipeline.draw_samples(pair_ids, 1024)).cuda()
for _ in range(3): pipe.run_device(ring, cnt, vox, voff, smp, pair_ids)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
acc = {}
for it in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0 = ev(); kpts, _kpix, n = ctx.select_keypoints(ring, cnt, None); e1 = ev()
    packed, _, _ = ctx.gather_patches(kpts, vox, voff, n); e2 = ev()
    feat = ctx.encode_frames(packed); e3 = ev()
    h1 = time.perf_counter()
    thr = torch.full((P,), 0.4, dtype=torch.float32, device="cuda")
    e4 = ev()
    host = torch.cat([result, rt], 1).cpu().numpy(); e5 = ev()
    h2 = time.perf_counter()
    failed = np.flatnonzero(host[:, 12] == 0)
    poses = np.zeros((P, 16), np.float32)
    if failed.size: pipe._ladder(failed, kpts, pair_idx, pair_ids, poses)
    torch.cuda.synchronize(); h3 = time.perf_counter()
    for k, v in (("select", e0.elapsed_time(e1)), ("gather", e1.elapsed_time(e2)), ("encode", e2.elapsed_time(e3)),
                 ("pairs", e3.elapsed_time(e4)), ("d2h", e4.elapsed_time(e5)), ("host_enqueue_frames_ms", (h1 - t0) * 1e3),
                 ("wall_to_d2h_ms", (h2 - t0) * 1e3), ("ladder_ms", (h3 - h2) * 1e3), ("n_failed", failed.size),
                 ("wall_total_ms", (h3 - t0) * 1e3)):
        acc.setdefault(k, []).append(v)
for k, v in acc.items(): print("%-24s %8.3f" % (k, float(np.mean(v))))
