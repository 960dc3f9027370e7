"""Host-side cost of one pipelined step (python + ctypes + torch allocator), measured while the device is busy:
python tools/host_overhead.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api, pipeline, synth

P = 32
ctx = api.default_context(); pipe = pipeline.OdometryPipeline(ctx)
d = synth.make_frames(P + 1, seed=1)
host = {k: torch.from_numpy(d[k]).pin_memory() for k in ("ring3", "counter", "vox")}
voff = d["vox_offsets"]; ids = list(range(P))
batch = ("rings", host["ring3"], host["counter"], host["vox"], voff, ids)
for _ in pipe.run_host_stream([batch] * 3): pass
acc = {"upload": [], "enqueue": [], "collect": []}
pending = None
up = pipe._upload(batch, 1)
torch.cuda.synchronize()
t_all = time.perf_counter()
for it in range(20):
    t0 = time.perf_counter(); nxt = pipe._upload(batch, 1)
    t1 = time.perf_counter(); h = pipe._enqueue(up)
    t2 = time.perf_counter()
    if pending is not None: pipe._collect(pending)
    t3 = time.perf_counter()
    pending, up = h, nxt
    acc["upload"].append(t1 - t0); acc["enqueue"].append(t2 - t1); acc["collect"].append(t3 - t2)
pipe._collect(pending)
torch.cuda.synchronize()
print("wall per step %.3f ms" % ((time.perf_counter() - t_all) / 20 * 1e3))
for k, v in acc.items(): print("%-8s host %.3f ms (median)" % (k, 1e3 * float(np.median(v))))
# per-kernel device time inside the streamed steps (H2D of the next batch running underneath) vs isolated steps
def per_kernel(fn, n):
    ctx.profile(True); ctx.profile_fetch(); fn(); prof = ctx.profile_fetch(); ctx.profile(False)
    return {k: v[1] / n for k, v in prof.items()}
def streamed():
    for _ in pipe.run_host_stream([batch] * 10): pass
dev = {k: torch.from_numpy(d[k]).cuda() for k in ("ring3", "counter", "vox")}
def isolated():
    for _ in range(10): pipe.run_device(dev["ring3"], dev["counter"], dev["vox"], voff, None, ids)
a, b = per_kernel(streamed, 10), per_kernel(isolated, 10)
print("kernel                         streamed  isolated (ms/step)")
for k in b: print("%-30s %8.4f  %8.4f" % (k, a.get(k, 0), b[k]))
print("%-30s %8.4f  %8.4f" % ("sum", sum(a.values()), sum(b.values())))
# finer: the calls inside _enqueue
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in pipe.run_host_stream([batch] * 10): pass
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
