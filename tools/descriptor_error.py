"""Measured descriptor error of the tensor-core encoder vs the torch-CPU fp32 oracle and vs the golden
Features/*.mat (TF 1.14): python tools/descriptor_error.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import golden_data as G
from caelo_b200 import api
from oracle import oracle

for tag in G.FRAMES[:2]:
    f, rr = G.frame(tag), G.refrun(tag)
    ref = G.unpack_patches(rr["patches_packed"])
    enc = api.load_model(api.WEIGHT_DIR + "/encoder.npz")
    got = api.GetFeaturesFromPatches(enc, ref)
    want = oracle.get_features_from_patches(ref)
    e = np.abs(got - want)
    eg = np.abs(got - f["golden_Features"]).max(1)
    print(tag, "vs oracle: max %.3e mean %.3e | per scale max" % (e.max(), e.mean()),
          [float(e[:, s * 20:(s + 1) * 20].max()) for s in range(3)],
          "| vs golden: rows<1e-5:", int((eg < 1e-5).sum()), "median row max %.2e" % np.median(eg))
rng = np.random.default_rng(3)
x = (rng.random((256, 16, 16, 16, 1)) < 0.3).astype(np.float32)
got, want = enc.predict(x), oracle.encoder_predict(x)
print("dense random patches: max %.3e mean %.3e" % (np.abs(got - want).max(), np.abs(got - want).mean()))
# elementwise view (north_star: "descriptors within 1e-4 relative"): which absolute floor does rtol = 1e-4 need?
for tag in G.FRAMES:
    f, rr = G.frame(tag), G.refrun(tag)
    ref = G.unpack_patches(rr["patches_packed"])
    got = api.GetFeaturesFromPatches(enc, ref)
    want = oracle.get_features_from_patches(ref)
    e = np.abs(got - want)
    for atol in (0.0, 1e-7, 1e-6, 2e-6, 5e-6, 1e-5):
        bad = e > 1e-4 * np.abs(want) + atol
        print(tag, "rtol 1e-4 atol %.0e: %d of %d elements outside" % (atol, int(bad.sum()), e.size))
    rel = e / np.maximum(np.abs(want), 1e-30)
    big = np.abs(want) > 1e-2
    print(tag, "max rel err where |ref| > 1e-2: %.3e; max abs err %.3e; p99.9 abs %.3e" % (rel[big].max(), e.max(), np.quantile(e, 0.999)))
