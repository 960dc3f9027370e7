"""A/B timing of the encoder kernels on one bench-shaped batch (33 frames x 3 x 1024 patches):
    python tools/encoder_ab.py            -> per-kernel ms for CAELO_CONV12_PAIR = 1 / 0 , CAELO_CONV3_OCT = 1 / 0 and with parts of the kernels switched off"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api, synth

ctx = api.default_context()
d = synth.make_frames(33, seed=1)
ring, cnt = api._dev(d["ring3"]), api._dev(d["counter"])
kpts, _, n = ctx.select_keypoints(ring, cnt, None)
packed, _, _ = ctx.gather_patches(kpts, api._dev(d["vox"]), d["vox_offsets"], n)
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
ref = None
for name, env in (("defaults (conv12 pair, conv3 oct)", {}), ("conv12 one patch", {"CAELO_CONV12_PAIR": "0"}),
                  ("conv12 pair, no bg skip", {"CAELO_CONV12_SKIP_BG": "0"}),
                  ("conv3 one patch (M=64)", {"CAELO_CONV3_OCT": "0"}),
                  ("conv3 oct, producers write nothing (wrong)", {"CAELO_CONV3_DBG": "1"}),
                  ("conv3 oct, epilogue only drains (wrong)", {"CAELO_CONV3_DBG": "2"}),
                  ("conv3 oct, no MMAs (wrong)", {"CAELO_CONV3_DBG": "4"}),
                  ("conv3 oct, MMAs only (wrong)", {"CAELO_CONV3_DBG": "3"}),
                  ("conv12 pair, no MMAs (wrong results)", {"CAELO_CONV12_DBG": "1"}),
                  ("conv12 pair, no conv1 pass 2 (wrong results)", {"CAELO_CONV12_DBG": "2"}),
                  ("conv12 pair, neither (wrong results)", {"CAELO_CONV12_DBG": "3"})):
    os.environ.update(env)
    for _ in range(3):
        feat = ctx.encode_frames(packed)
    ctx.profile(True); ctx.profile_fetch()
    for _ in range(10):
        flush.zero_()
        feat = ctx.encode_frames(packed)
    torch.cuda.synchronize()
    prof = ctx.profile_fetch(); ctx.profile(False)
    for k in env: del os.environ[k]
    f = feat.cpu().numpy()
    if ref is None: ref = f
    print("%-30s %s  max|d - default| %.2e" % (name, {k: round(v[1] / v[0], 4) for k, v in prof.items()}, float(np.abs(f - ref).max())))
