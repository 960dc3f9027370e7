"""One batched step (33 synthetic frames, 32 pairs) between cudaProfilerStart/Stop, for ncu:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python tools/ncu_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on -o R python tools/ncu_step.py [--scans]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api, pipeline, synth

P = 32
ctx = api.default_context(); pipe = pipeline.OdometryPipeline(ctx)
d = synth.make_frames(P + 1, seed=1)
ids = list(range(P))
if "--scans" in sys.argv:
    soff = np.zeros(P + 2, np.int64); soff[1:] = np.cumsum([s.shape[0] for s in d["scans"]])
    pts = torch.from_numpy(np.concatenate(d["scans"], 0)).cuda()
    step = lambda: pipe.run_device_scans(pts, soff, None, ids)
else:
    ring, cnt, vox = (torch.from_numpy(d[k]).cuda() for k in ("ring3", "counter", "vox"))
    step = lambda: pipe.run_device(ring, cnt, vox, d["vox_offsets"], None, ids)
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
