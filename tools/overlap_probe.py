"""How much of the a6 stage (brick index + gather) and of the key-point selection hides under the encoder of ANOTHER batch
when they run on a second (lower-priority) stream: python tools/overlap_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api, synth

ctx = api.default_context()
d = synth.make_frames(33, seed=1)
ring, cnt, vox = api._dev(d["ring3"]), api._dev(d["counter"]), api._dev(d["vox"])
voff = d["vox_offsets"]
kpts, _, n = ctx.select_keypoints(ring, cnt, None)
ctx.bricks_build(vox, voff)
packed = ctx.bricks_gather(kpts, n)
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
main = torch.cuda.Stream(priority=-1)
side = torch.cuda.Stream(priority=0)

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    torch.cuda.current_stream().wait_stream(main); torch.cuda.current_stream().wait_stream(side)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def enc():
    with torch.cuda.stream(main): ctx.encode_frames(packed)
def a6(stream):
    with torch.cuda.stream(stream):
        ctx.bricks_build(vox, voff); ctx.bricks_gather(kpts, n)
def sel(stream):
    with torch.cuda.stream(stream): ctx.select_keypoints(ring, cnt, None)
def both(fn):
    main.wait_stream(torch.cuda.current_stream()); side.wait_stream(torch.cuda.current_stream())
    fn()

t_enc = timed(lambda: both(enc))
t_a6 = timed(lambda: both(lambda: a6(main)))
t_sel = timed(lambda: both(lambda: sel(main)))
print("alone: encoder %.3f  a6 %.3f  select %.3f ms" % (t_enc, t_a6, t_sel))
print("encoder + a6 on one stream        %.3f" % timed(lambda: both(lambda: (a6(main), enc()))))
print("encoder (hi prio) || a6 (side)    %.3f" % timed(lambda: both(lambda: (enc(), a6(side)))))
print("a6 (side) queued first || encoder %.3f" % timed(lambda: both(lambda: (a6(side), enc()))))
print("encoder || a6 + select (side)     %.3f  (serial %.3f)" % (timed(lambda: both(lambda: (enc(), a6(side), sel(side)))), t_enc + t_a6 + t_sel))
print("encoder || select (side)          %.3f  (serial %.3f)" % (timed(lambda: both(lambda: (enc(), sel(side)))), t_enc + t_sel))
