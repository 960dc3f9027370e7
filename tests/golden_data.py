"""Loaders for the committed fixtures under tests/golden/ (see make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAMES = ["00_000000", "00_000001", "01_000495", "01_000496"]
PAIRS = {"00": ("00_000000", "00_000001"), "01": ("01_000495", "01_000496")}
ImgH, ImgW = 69, 1800

_cache = {}


def frame(tag):
    """dict(ring5 (69,1800,5) f32, counter (69,1800) i32, vox0/1/2 int16, golden_KeyPts, golden_Features)."""
    if tag in _cache:
        return _cache[tag]
    z = np.load(os.path.join(GOLDEN, "frame_%s.npz" % tag))
    ring = np.zeros((ImgH * ImgW, 5), np.float32)
    ring[z["ring_idx"]] = z["ring_val"]
    counter = np.zeros(ImgH * ImgW, np.int32)
    counter[z["ring_idx"]] = z["counter_val"]
    d = dict(ring5=ring.reshape(ImgH, ImgW, 5), counter=counter.reshape(ImgH, ImgW),
             vox0=z["vox0"], vox1=z["vox1"], vox2=z["vox2"],
             golden_KeyPts=z["golden_KeyPts"], golden_Features=z["golden_Features"])
    d["ring3"] = np.ascontiguousarray(d["ring5"][0:64, 0:1792, 0:3])
    d["counter_i8"] = d["counter"].astype(np.int8)
    _cache[tag] = d
    return d


def refrun(tag):
    z = np.load(os.path.join(GOLDEN, "refrun_%s.npz" % tag))
    return {k: z[k] for k in z.files}


def pose(seq):
    z = np.load(os.path.join(GOLDEN, "pose_%s.npz" % seq))
    return {k: z[k] for k in z.files}


def usip(seq):
    z = np.load(os.path.join(GOLDEN, "usip_%s.npz" % seq))
    return {k: z[k] for k in z.files}


def unpack_patches(packed):
    """(3,K,512) uint8 -> list of 3 (K,16,16,16,1) float32."""
    out = []
    for s in range(packed.shape[0]):
        bits = np.unpackbits(packed[s], axis=1).astype(np.float32)
        out.append(bits.reshape(-1, 16, 16, 16, 1))
    return out


SCANS = ["00_000000", "01_000495"]


def scan(tag):
    """Raw DemoData scan (N,4) f32 + what the unmodified reference Voxelization returned for it."""
    z = np.load(os.path.join(GOLDEN, "scan_%s.npz" % tag))
    return {k: z[k] for k in z.files}
