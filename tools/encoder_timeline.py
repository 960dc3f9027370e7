"""Dump per-patch phase timings (SM cycles) of the encoder's conv12 kernel: python tools/encoder_timeline.py"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caelo_b200 import api, synth

ctx = api.default_context()
d = synth.make_frames(2, seed=1)
ring, cnt = api._dev(d["ring3"]), api._dev(d["counter"])
kpts, _, n = ctx.select_keypoints(ring, cnt, None)
packed, _, _ = ctx.gather_patches(kpts, api._dev(d["vox"]), d["vox_offsets"], n)
packed = packed.repeat(8, 1, 1, 1).contiguous()          # 49k patches
grid = 2 * ctx.lib.caelo_num_sms(ctx.h)
tl_all = torch.zeros((grid * 64 * 16 + 1024 * 8,), dtype=torch.int64, device="cuda")
ctx.encode_frames(packed)
ctx.check(ctx.lib.caelo_debug_set_timeline(ctx.h, ctypes.c_void_p(tl_all.data_ptr())))
ctx.encode_frames(packed)
torch.cuda.synchronize()
ctx.lib.caelo_debug_set_timeline(ctx.h, None)
tl = tl_all[:grid * 64 * 16].view(grid, 64, 16)
dn = tl_all[grid * 64 * 16:].view(-1, 8).cpu().numpy().astype(np.float64)
dn = dn[dn[:, 0] > 0]
print("dense_tc per-CTA cycles: start->accum-done %.0f  W2 fill+Hs fill %.0f  dense2 %.0f  total %.0f (n=%d)" % ((dn[:,3]-dn[:,0]).mean(), (dn[:,4]-dn[:,3]).mean(), (dn[:,5]-dn[:,4]).mean(), (dn[:,5]-dn[:,0]).mean(), len(dn)))
t = tl.cpu().numpy()
t = t[t[:, 10, 0] > 0][:, 4:60, :].astype(np.float64)          # CTAs that ran (the pair kernel launches one per SM)
print("CTAs with stamps:", t.shape[0])
def stat(x): return "mean %8.0f  p50 %8.0f  p90 %8.0f" % (x.mean(), np.median(x), np.percentile(x, 90))
print("conv1            1-0 :", stat(t[..., 1] - t[..., 0]))
print("  restore+stage  2-0 :", stat(t[..., 2] - t[..., 0]))
print("  pass 1         7-2 :", stat(t[..., 7] - t[..., 2]))
print("  pass 2         8-7 :", stat(t[..., 8] - t[..., 7]))
print("  fence + arrive 1-8 :", stat(t[..., 1] - t[..., 8]))

if t[..., 13].max() > 0:
    sel = t[..., 14] > 0
    print("  pass 2, warp 0's first round (pairs with listed cells: %.0f %%, mean n %.0f):" % (100 * sel.mean(), t[..., 14][sel].mean()))
    for a, b, name in ((7, 9, "list entry + row loads"), (9, 10, "window (nibbles + 6 shuffles)"), (10, 11, "3 table rows + adds"),
                       (11, 12, "max over 8 lanes (7 shuffles)"), (12, 13, "tanh + split + stores"), (13, 8, "later rounds")):
        print("    %-32s %s" % (name, stat((t[..., b] - t[..., a])[sel])))
    for lo, hi in ((1, 64), (65, 128), (129, 256), (257, 1024)):
        m = (t[..., 14] >= lo) & (t[..., 14] <= hi)
        if m.any(): print("    n in [%4d, %4d]: %4.1f %% of pairs, pass 2 mean %6.0f" % (lo, hi, 100 * m.mean(), (t[..., 8] - t[..., 7])[m].mean()))
if t[..., 9].max() > 0:   # epilogue-warp role of the pair kernel: stamps of step j are in row j (restore, listing), of its epilogue in row j-2
    e = t[:, 2:-2, :]
    print("epilogue warps, step j: tfull(j-2) -> restore done      :", stat(e[..., 9] - t[:, :-4, 3]))
    print("                        restore -> cells listed         :", stat(e[..., 10] - e[..., 9]))
    print("                        listed -> fence + arrive        :", stat(e[..., 2] - e[..., 10]))
    print("                        epilogue proper of pair j-2     :", stat(t[:, :-4, 4] - t[:, :-4, 11]))
print("mma issue        6-5 :", stat(t[..., 6] - t[..., 5]))
print("mbar wait        3-1 :", stat(t[..., 3] - t[..., 1]))
print("epilogue         4-3 :", stat(t[..., 4] - t[..., 3]))

print("iteration  next0-0 :", stat(t[:, 1:, 0] - t[:, :-1, 0]))
for r in range(3):
    sel = [i for i in range(t.shape[1] - 1) if (i + 4) % 3 == r]
    print("  iterations with i %% 3 == %d: iteration %s | pass1 %6.0f pass2 %6.0f mma %6.0f epi %6.0f" % (
        r, stat((t[:, 1:, 0] - t[:, :-1, 0])[:, sel]), (t[..., 7] - t[..., 2])[:, sel].mean(), (t[..., 8] - t[..., 7])[:, sel].mean(),
        (t[..., 6] - t[..., 5])[:, sel].mean(), (t[..., 4] - t[..., 3])[:, sel].mean()))
raw = tl.cpu().numpy()
for cta in (0, 100):
    print("CTA", cta, "(cycles relative to iteration-10 start; columns = stamps 0..7)")
    base = raw[cta, 10, 0]
    for i in range(10, 15):
        print("  it %2d:" % i, " ".join("%8d" % (raw[cta, i, s] - base) for s in range(9)))
