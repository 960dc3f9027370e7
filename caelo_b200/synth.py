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


# Beam pattern of the sensor KITTI was recorded with (Velodyne HDL-64E): 64 lasers in two blocks — the upper 32 are
# 1/3 degree apart, the lower 32 are 1/2 degree apart — firing 2083 times per revolution (0.1728 degrees).  The ring
# image has 0.2-degree columns and 0.425-degree rows, so ~16 % of a beam's shots share a column with the previous
# shot and neighbouring upper beams share rows: ~120 k points per scan, ~25-30 k pixels hit by more than one point
# (124,668 points, 27,959 multi-hit pixels on demo frame 00/000000) — the load that
# ProjectPC2SphericalRing's last-writer-wins rule and Voxelization's duplicate handling see on real data.
N_AZIMUTH = 2083
BEAM_ELEVATION_DEG = np.r_[1.9 - np.arange(32) / 3.0, -8.83 - np.arange(32) / 2.0]
DROPOUT = 0.06


def _rays():
    el = BEAM_ELEVATION_DEG * _DEG
    az = math.pi - (2 * math.pi / N_AZIMUTH) * (np.arange(N_AZIMUTH) + 0.5)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d = np.stack([ce * np.cos(az)[None], ce * np.sin(az)[None], np.broadcast_to(se, (64, N_AZIMUTH))], -1)
    return d.reshape(-1, 3).astype(np.float32)      # file order: beam by beam, each a full revolution


_RAYS = {}


def _rays_on(device):
    import torch
    key = str(device)
    if key not in _RAYS:
        _RAYS[key] = torch.from_numpy(_rays()).to(device)
    return _RAYS[key]


def sensor_pose(frame: int, step: float = 0.7, yaw_step_deg: float = 0.3):
    """World pose of the sensor at ``frame``: x_world = Rz(yaw) x_sensor + pos.  ``step`` metres forward and
    ``yaw_step_deg`` of heading per frame, weaving +-2 m across the road (2 cm per frame at the start) — bounded, so a
    4541-frame drive stays on the road."""
    yaw = math.radians(yaw_step_deg) * frame
    cy, sy = math.cos(yaw), math.sin(yaw)
    return (np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], np.float32),
            np.array([step * frame, 2.0 * math.sin(0.01 * frame), 0.0], np.float32))


def scan_torch(world: World, frame: int, seed: int, device="cpu", step: float = 0.7, yaw_step_deg: float = 0.3):
    """One LiDAR scan as a torch tensor [N,4] f32 (x, y, z in the sensor frame, intensity) on ``device``: every ray
    of the beam pattern is cast against the ground plane and the boxes of the tiles in reach (slab test), returns
    outside [3, 80] m and a random 6 % are dropped, 2 cm range noise.  The random numbers come from numpy
    (``default_rng(seed)``) whatever the device, so the CPU and the GPU produce the same scan up to the rounding
    of the ray casting."""
    import torch
    rays = _rays_on(device)
    Rw, pos = sensor_pose(frame, step, yaw_step_deg)
    rng = np.random.default_rng(seed)
    nr = rays.shape[0]
    u = torch.from_numpy(rng.random(nr, dtype=np.float32)).to(device)
    noise = torch.from_numpy(rng.standard_normal(nr, dtype=np.float32) * np.float32(0.02)).to(device)
    inten = torch.from_numpy(rng.random(nr, dtype=np.float32)).to(device)
    pos_t = torch.from_numpy(pos).to(device)
    d = rays @ torch.from_numpy(np.ascontiguousarray(Rw.T)).to(device)
    t = torch.full((nr,), float("inf"), dtype=torch.float32, device=device)
    down = d[:, 2] < -1e-6
    t = torch.where(down, (-1.73 - pos_t[2]) / d[:, 2], t)
    inv = 1.0 / d
    wlo, whi = world.boxes_near(float(pos[0]))
    lo_t, hi_t = torch.from_numpy(wlo).to(device), torch.from_numpy(whi).to(device)
    for b0 in range(0, lo_t.shape[0], 32):
        t1 = (lo_t[None, b0:b0 + 32] - pos_t) * inv[:, None, :]
        t2 = (hi_t[None, b0:b0 + 32] - pos_t) * inv[:, None, :]
        tn = torch.minimum(t1, t2).amax(-1)
        tf = torch.maximum(t1, t2).amin(-1)
        hit = (tf >= tn) & (tn > 0)
        t = torch.minimum(t, torch.where(hit, tn, torch.full_like(tn, float("inf"))).amin(1))
    ok = torch.isfinite(t) & (t >= 3.0) & (t <= 80.0) & (u > DROPOUT)
    pts = rays[ok] * (t[ok] + noise[ok])[:, None]
    return torch.cat([pts, inten[ok][:, None]], 1)


def scan(world: World, frame: int, seed: int, step: float = 0.7, yaw_step_deg: float = 0.3):
    """One LiDAR scan (N,4) f32 in the sensor frame; sensor moves ``step`` m forward per frame."""
    return scan_torch(world, frame, seed, "cpu", step, yaw_step_deg).numpy()


def make_scans(n_frames: int, seed: int = 0, first_frame: int = 0, device="cpu"):
    """Frames first_frame .. first_frame+n_frames-1 of the drive through World(seed) -> (pts [sumN,4] f32 tensor on
    ``device``, row offsets int64 [F+1]).  On a GPU a 4541-frame sequence (configs[2]) takes seconds."""
    import torch
    world = World(seed)
    parts = [scan_torch(world, first_frame + f, seed * 100003 + first_frame + f, device) for f in range(n_frames)]
    off = np.zeros(n_frames + 1, np.int64)
    off[1:] = np.cumsum([p.shape[0] for p in parts])
    return torch.cat(parts, 0), off


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
    ring = np.zeros((ImgH * ImgW, 5), np.float32)
    flat = row * ImgW + col
    counter = np.bincount(flat, minlength=ImgH * ImgW).astype(np.int32).reshape(ImgH, ImgW)
    # the LAST point in file order wins a pixel (SphericalRing.py:91-92): first occurrence in the reversed order
    pix, first_rev = np.unique(flat[::-1], return_index=True)
    last = flat.shape[0] - 1 - first_rev
    ring[pix, 0:4] = pc[last, 0:4]
    ring[pix, 4] = r[last]
    return ring.reshape(ImgH, ImgW, 5), counter


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


def make_frames(n_frames: int, seed: int = 0, first_frame: int = 0, device=None):
    """-> dict(ring3 [F,64,1792,3] f32, counter [F,69,1800] i8, vox int16 [sumV,3], vox_offsets int64 [3F+1],
    scans list).  ``device``: where the rays are cast (default: the GPU when there is one)."""
    world = World(seed)
    if device is None:
        import torch
        device = "cuda" if torch.cuda.is_available() else "cpu"
    ring3 = np.zeros((n_frames, 64, 1792, 3), np.float32)
    counter = np.zeros((n_frames, ImgH, ImgW), np.int8)
    vox, off = [], [0]
    npts = []
    scans = [scan_torch(world, first_frame + f, seed * 100003 + first_frame + f, device).cpu().numpy()
             for f in range(n_frames)]
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
