"""Where does the from-scans step spend its stream time?  Event-times each stage of scans_to_descriptors, 10 calls queued."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api, pipeline, synth
ctx = api.default_context(); pipe = pipeline.OdometryPipeline(ctx)
P = 32
d = synth.make_frames(P + 1, seed=1)
soff = np.zeros(P + 2, np.int64); soff[1:] = np.cumsum([s.shape[0] for s in d["scans"]])
pts = torch.from_numpy(np.concatenate(d["scans"], 0)).cuda()
ids = list(range(P))
def timed(name, fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print("%-28s device %.3f ms/call   host enqueue %.3f ms/call" % (name, a.elapsed_time(b) / n, (t1 - t0) * 1e3 / n), flush=True)
r = ctx.project_ring(pts, soff, want=("ring3", "counter_i8"))
kpts, _px, n = ctx.select_keypoints(r["ring3"], r["counter_i8"], None)
timed("project_ring", lambda: ctx.project_ring(pts, soff, want=("ring3", "counter_i8")))
timed("select_keypoints", lambda: ctx.select_keypoints(r["ring3"], r["counter_i8"], None))
timed("bricks_build_scans", lambda: ctx.bricks_build_scans(pts, soff))
timed("gather_patches_scans", lambda: ctx.gather_patches_scans(kpts, pts, soff, n))
packed = ctx.gather_patches_scans(kpts, pts, soff, n)[0]
timed("encode_frames", lambda: ctx.encode_frames(packed))
timed("draw_samples", lambda: ctx.draw_samples(ids, 1024, rounds=3))
timed("scans_to_descriptors", lambda: pipe.scans_to_descriptors(pts, soff))
timed("enqueue_device_scans+collect", lambda: pipe.collect(pipe.enqueue_device_scans(pts, soff, None, ids)))
hs = []
def q(): hs.append(pipe.enqueue_device_scans(pts, soff, None, ids))
timed("enqueue_device_scans (async)", q)
for h in hs: pipe.collect(h)
ring, cnt, vox = (torch.from_numpy(d[k]).cuda() for k in ("ring3", "counter", "vox"))
timed("rings: enqueue+collect", lambda: pipe.collect(pipe.enqueue_device(ring, cnt, vox, d["vox_offsets"], None, ids)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
hs = [pipe.enqueue_device_scans(pts, soff, None, ids) for _ in range(10)]
pr.disable()
for h in hs: pipe.collect(h)
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
