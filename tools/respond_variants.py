"""A/B timing of the fused respond + score kernel variants (CAELO_RESPOND_VARIANT, read once per process):
    for v in 0 1 2 3; do CAELO_RESPOND_VARIANT=$v python tools/respond_variants.py; done"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api, synth
ctx = api.default_context()
d = synth.make_frames(33, seed=1)
ring, cnt = torch.from_numpy(d["ring3"]).cuda(), torch.from_numpy(d["counter"]).cuda()
for _ in range(3): ctx.select_keypoints(ring, cnt, None)
ctx.profile(True); ctx.profile_fetch()
for _ in range(20): ctx.select_keypoints(ring, cnt, None)
prof = ctx.profile_fetch()
print("variant", os.environ.get("CAELO_RESPOND_VARIANT", "default"), {k: round(v[1] / v[0], 4) for k, v in prof.items()})
