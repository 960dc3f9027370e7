"""Seeded synthetic KITTI-seq-00-shaped inputs for the hot path: a 64-beam,
1800-azimuth scan of a ground plane with boxes and poles, projected to the 69x1800 spherical
ring and voxelised at the three scales — i.e. exactly what the reference's offline stages
(BatchPreprocess option 1, BatchVoxelization) hand to the hot path.  numpy only; used by
bench.py, smoke tests and the multi-rank tests.  There is no network for real KITTI data."""
from __future__ import annotations

import math

import numpy as np

# SphericalRing.py:28-57
_DEG = math.pi / 180
AzimuthResolution = 0.20 * _DEG
VerticalViewDown = -24.8 * _DEG
VerticalViewUp = 2.0 * _DEG
VerticalResolution = (VerticalViewUp - VerticalViewDown) / 63
VerticalPixelsOffset = -VerticalViewDown / VerticalResolution
ImgH, ImgW = 69, 1800
# Voxel.py:40-52
VIS = np.array([156 / 2 * 1.28, 156 / 2 * 1.28, 23 / 2 * 1.28])
VSIZES = [0.02, 0.02 * 8, 0.02 * 32]


class World:
    """A street scene that repeats every ``PERIOD`` metres along the driving direction, so that every frame of an
    arbitrarily long drive (rank r of an 8-GPU run starts 0.7 * 32 * r metres down the road) sees the same kind of
    surroundings: ~50 boxes and ~40 poles per tile on a ground plane, side walls along the road and cross walls
    with a gap for the road where two tiles meet (the upper beams return too: KITTI seq 00 is urban, ~88k of 115k
    pixels hit)."""
    W, F, Hh = 58.0, 76.0, 25.0
    PERIOD = 58.0 + 76.0

    def __init__(self, seed: int, n_boxes: int = 50, n_poles: int = 40):
        rng = np.random.default_rng(seed)
        c = rng.uniform(-70, 70, (n_boxes, 2))
        c[:, 0] += 20
        half = rng.uniform(1.0, 6.0, (n_boxes, 2))
        h = rng.uniform(2.0, 12.0, n_boxes)
        pc = rng.uniform(-50, 50, (n_poles, 2))
        pc[:, 0] += 15
        ph = np.full((n_poles, 2), 0.15)
        hh = rng.uniform(3.0, 9.0, n_poles)
        cen = np.concatenate([c, pc])
        hal = np.concatenate([half, ph])
        hei = np.concatenate([h, hh])
        keep = (np.abs(cen[:, 0]) > 8) | (np.abs(cen[:, 1]) > 4)   # keep the start of the road clear
        keep &= np.abs(cen[:, 1]) - hal[:, 1] > 3.0                # ... and the road itself, all along the tile
        lo = np.c_[cen - hal, np.full(len(cen), -1.73)][keep]
        hi = np.c_[cen + hal, hei - 1.73][keep]
        W, F, Hh, G = self.W, self.F, self.Hh, 6.0                 # G: half width of the road gap in the cross walls
        walls_lo = [[F, -W, -1.73], [F, G, -1.73], [-W, -W - 1, -1.73], [-W, W, -1.73]]
        walls_hi = [[F + 1, -G, Hh], [F + 1, W, Hh], [F, -W, Hh], [F, W + 1, Hh]]
        self.lo = np.concatenate([lo, walls_lo]).astype(np.float32)
        self.hi = np.concatenate([hi, walls_hi]).astype(np.float32)

    def boxes_near(self, x: float, reach: float = 85.0):
        """Boxes (lo, hi) of the tiles around position x along the road that can be hit within ``reach`` metres."""
        k0 = int(np.floor((x + self.W) / self.PERIOD))
        los, his = [], []
        for k in (k0 - 1, k0, k0 + 1):
            off = np.array([k * self.PERIOD, 0.0, 0.0], np.float32)
            lo, hi = self.lo + off, self.hi + off
            near = (hi[:, 0] > x - reach) & (lo[:, 0] < x + reach)
            los.append(lo[near])
            his.append(hi[near])
        return np.concatenate(los), np.concatenate(his)


def _rays():
    el = VerticalViewDown + VerticalResolution * (np.arange(64) + 0.5)   # beam centred in its ring row
    az = math.pi - AzimuthResolution * (np.arange(1800) + 0.5)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d = np.stack([ce * np.cos(az)[None], ce * np.sin(az)[None], np.broadcast_to(se, (64, 1800))], -1)
    return d.reshape(-1, 3).astype(np.float32)


_RAYS = None


def scan(world: World, frame: int, seed: int, step: float = 0.7, yaw_step_deg: float = 0.3):
    """One LiDAR scan (N,4) f32 in the sensor frame; sensor moves ``step`` m forward per frame."""
    global _RAYS
    if _RAYS is None:
        _RAYS = _rays()
    rng = np.random.default_rng(seed)
    yaw = math.radians(yaw_step_deg) * frame
    pos = np.array([step * frame, 0.02 * frame, 0.0], np.float32)
    cy, sy = math.cos(yaw), math.sin(yaw)
    Rw = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], np.float32)
    d = _RAYS @ Rw.T
    t = np.full(d.shape[0], np.inf, np.float32)
    down = d[:, 2] < -1e-6
    t[down] = (-1.73 - pos[2]) / d[down, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (1.0 / d).astype(np.float32)
        wlo, whi = world.boxes_near(float(pos[0]))
        for b0 in range(0, wlo.shape[0], 16):
            lo = wlo[b0:b0 + 16][None]
            hi = whi[b0:b0 + 16][None]
            t1 = (lo - pos) * inv[:, None, :]
            t2 = (hi - pos) * inv[:, None, :]
            tn = np.minimum(t1, t2).max(-1)
            tf = np.maximum(t1, t2).min(-1)
            hit = (tf >= tn) & (tn > 0)
            tb = np.where(hit, tn, np.inf).min(1)
            t = np.minimum(t, tb)
    ok = np.isfinite(t) & (t >= 3.0) & (t <= 80.0) & (rng.random(t.shape[0]) > 0.25)
    rng_t = t[ok] + rng.normal(0, 0.02, int(ok.sum())).astype(np.float32)   # 2 cm range noise
    pts = (_RAYS[ok] * rng_t[:, None]).astype(np.float32)                    # sensor frame
    inten = rng.random((pts.shape[0], 1)).astype(np.float32)
    return np.concatenate([pts, inten], 1)


def project_ring(pc: np.ndarray):
    """Vectorised restatement of ProjectPC2SphericalRing (SphericalRing.py:72-94) for synthetic
    input: float64 angles, last point in file order wins a pixel.  -> ring (69,1800,5), counter."""
    xyz = pc[:, :3]
    r = np.linalg.norm(xyz, axis=1)
    keep = r > 0
    pc, xyz, r = pc[keep], xyz[keep], r[keep]
    col = ((math.pi - np.arctan2(xyz[:, 1].astype(np.float64), xyz[:, 0].astype(np.float64))) / AzimuthResolution).astype(np.int64)
    beta = np.arcsin((xyz[:, 2] / r).astype(np.float64))
    row = ImgH - (beta / VerticalResolution + VerticalPixelsOffset).astype(np.int64)
    ok = (row >= 0) & (row < ImgH) & (col >= 0) & (col < ImgW)
    row, col, pc, r = row[ok], col[ok], pc[ok], r[ok]
    ring = np.zeros((ImgH, ImgW, 5), np.float32)
    counter = np.zeros((ImgH, ImgW), np.int32)
    ring[row, col, 0:4] = pc[:, 0:4]
    ring[row, col, 4] = r
    np.add.at(counter, (row, col), 1)
    return ring, counter


def voxelize(pc: np.ndarray):
    """Three occupied-voxel lists (int16 (V,3)) as Voxelization (Voxel.py:100-173) defines them: float64
    (p + Visible) / size truncated, the 2 cm index through the 1.28 m block as the reference computes it
    (:120-139); unique, in first-seen order (the reference additionally groups list 0 by block — an order
    the hot path never sees)."""
    p = pc[:, :3].astype(np.float64)
    ok = (np.abs(p[:, 0]) <= VIS[0]) & (np.abs(p[:, 1]) <= VIS[1]) & (np.abs(p[:, 2]) <= VIS[2])
    p = p[ok] + VIS
    out = []
    for i, s in enumerate(VSIZES):
        if i == 0:
            b = (p / 1.28).astype(np.int64)
            v = ((p - b * 1.28) / s).astype(np.int64) + b * 64
        else:
            v = (p / s).astype(np.int64)
        key = (v[:, 0] << 40) | (v[:, 1] << 20) | v[:, 2]
        _, first = np.unique(key, return_index=True)
        out.append(v[np.sort(first)].astype(np.int16))
    return out


def make_frames(n_frames: int, seed: int = 0, first_frame: int = 0):
    """-> dict(ring3 [F,64,1792,3] f32, counter [F,69,1800] i8, vox int16 [sumV,3], vox_offsets int64 [3F+1],
    scans list)."""
    world = World(seed)
    ring3 = np.zeros((n_frames, 64, 1792, 3), np.float32)
    counter = np.zeros((n_frames, ImgH, ImgW), np.int8)
    vox, off = [], [0]
    npts = []
    from concurrent.futures import ThreadPoolExecutor
    import os
    # one generator per rank runs at the same time on the box: share the host cores
    n_ranks = max(1, int(os.environ.get("WORLD_SIZE", "1")))
    with ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 1) // n_ranks))) as ex:
        scans = list(ex.map(lambda f: scan(world, first_frame + f, seed * 100003 + first_frame + f),
                            range(n_frames)))
    for f in range(n_frames):
        pc = scans[f]
        ring, cnt = project_ring(pc)
        ring3[f] = ring[0:64, 0:1792, 0:3]
        counter[f] = np.minimum(cnt, 127).astype(np.int8)
        for v in voxelize(pc):
            vox.append(v)
            off.append(off[-1] + v.shape[0])
        npts.append(pc.shape[0])
    return dict(ring3=ring3, counter=counter, vox=np.concatenate(vox, 0), vox_offsets=np.asarray(off, np.int64),
                n_points=npts, scans=scans)
