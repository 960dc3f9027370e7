"""Host-side mirror of the reference's call surface for the odometry hot path.

Same names, argument meaning and return values as the reference functions, numpy in / numpy
out; underneath every call goes through the C ABI of libcaelo_b200.so (include/caelo.h) with
torch tensors used only as device buffers.  There is no CPU fallback.

    reference symbol (file:line)                          here
    keras.models.load_model (Match.py:313,324)            load_model -> B200Model
    model.predict (SphericalRing.py:407, Match.py:131)    B200Model.predict
    ProjectPC2SphericalRing (SphericalRing.py:72)         ProjectPC2SphericalRing
    Voxelization (Voxel.py:100)                           Voxelization
    GetKeyPtsByAE (SphericalRing.py:113)                  GetKeyPtsByAE
    GetKeyPtsFromRawFileName (SphericalRing.py:389)       GetKeyPtsFromRawFileName
    ExtendKeyPtsInShpericalRing (SphericalRing.py:294)    ExtendKeyPtsInShpericalRing
    GetPatchesList (Voxel.py:177)                         GetPatchesList
    GetFeaturesFromPatches (Match.py:130)                 GetFeaturesFromPatches
    SolveRT / RANSAC4RT / SolveRelativePose (Match.py)    SolveRT / RANSAC4RT / SolveRelativePose
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from .h5weights import layer_summary, read_keras_weights, read_model_config

_HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHT_DIR = os.path.join(_HERE, "weights")

# module-level constants the reference's drivers pull in with ``from X import *``
# (SphericalRing.py:28-57, Voxel.py:15-52)
nLines = 64
ImgH = 69
ImgW = 1800
CropWidth_SphericalRing = 8
Channels4AE = [0, 1, 2]
PatchSize = 16
Scales = 3
VoxelSize = 0.02
VoxelSizes = [VoxelSize, VoxelSize * 8, VoxelSize * 32]
VisibleLength = 156 / 2 * 1.28
VisibleWidth = 156 / 2 * 1.28
VisibleHeight = 23 / 2 * 1.28
nFixedKeyPts = 1024
MAX_TRIALS = 500


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class Context:
    """One per GPU: owns the caelo_ctx, the network weights and the device scratch."""

    def __init__(self, device: int = 0, respond_weights=None, encoder_weights=None):
        if not torch.cuda.is_available():
            raise _lib.CaeloError("no CUDA device: caelo_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)  # make torch own the primary context first
        h = ctypes.c_void_p()
        _lib.check(self.lib.caelo_create(device, ctypes.byref(h)), None, "caelo_create")
        self.h = h
        self.has_respond = False
        self.has_encoder = False
        self.set_respond_weights(respond_weights or _default_weights("respond"))
        self.set_encoder_weights(encoder_weights or _default_weights("encoder"))

    def check(self, rc, what=""):
        _lib.check(rc, self.h, what)

    def set_respond_weights(self, w):
        a = [np.ascontiguousarray(w[k], np.float32) for k in
             ("conv2d_1/kernel:0", "conv2d_1/bias:0", "conv2d_2/kernel:0", "conv2d_2/bias:0")]
        assert a[0].shape == (3, 3, 3, 32) and a[2].size == 256
        self.check(self.lib.caelo_set_respond_weights(self.h, *[_np_ptr(x) for x in a]), "set_respond_weights")
        self.has_respond = True
        self.__dict__.setdefault("_weight_owner", {}).pop("respond", None)      # see B200Model._bind

    def set_encoder_weights(self, w):
        keys = ("conv3d_1/kernel:0", "conv3d_1/bias:0", "conv3d_2/kernel:0", "conv3d_2/bias:0",
                "conv3d_3/kernel:0", "conv3d_3/bias:0", "dense_1/kernel:0", "dense_1/bias:0",
                "dense_2/kernel:0", "dense_2/bias:0")
        a = [np.ascontiguousarray(w[k], np.float32) for k in keys]
        assert a[0].shape == (3, 3, 3, 1, 8) and a[6].shape == (2048, 200) and a[8].shape == (200, 20)
        self.check(self.lib.caelo_set_encoder_weights(self.h, *[_np_ptr(x) for x in a]), "set_encoder_weights")
        self.has_encoder = True
        self.__dict__.setdefault("_weight_owner", {}).pop("encoder", None)

    @property
    def launches(self) -> int:
        return int(self.lib.caelo_launch_count(self.h))

    def profile(self, on: bool):
        self.check(self.lib.caelo_profile_enable(self.h, 1 if on else 0), "caelo_profile_enable")

    def profile_fetch(self):
        """{kernel name: (launches, total_ms)} since the last fetch (synchronises)."""
        buf = ctypes.create_string_buffer(1 << 16)
        self.check(self.lib.caelo_profile_fetch(self.h, buf, len(buf)), "caelo_profile_fetch")
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.rsplit(" ", 2)
            out[name] = (int(n), float(ms))
        return out

    def _scratch(self, name: str, shape, dtype) -> torch.Tensor:
        """A persistent device buffer for a TRANSIENT result of the batched pipeline (ring images, packed patches):
        the same tensor is handed out call after call.  Safe because every kernel that touches it is queued on the one
        compute stream; it spares the hot path its large torch allocations — with several batches queued ahead the
        caching allocator kept carving the freed 45-52 MB blocks up for the small per-batch results and went back to
        cudaMalloc (which synchronises the device) for the big ones: 6 ms of host time per step."""
        if not hasattr(self, "_scratch_bufs"):
            self._scratch_bufs = {}
        key = (name, tuple(shape), dtype)
        buf = self._scratch_bufs.get(key)
        if buf is None:
            if len(self._scratch_bufs) > 32:                      # many different batch shapes: start over
                self._scratch_bufs.clear()
            buf = self._scratch_bufs[key] = torch.empty(tuple(shape), dtype=dtype, device=self.device)
        return buf

    def close(self):
        if getattr(self, "h", None):
            self.lib.caelo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device-level wrappers (torch tensors in/out, asynchronous) ----------------------
    def respond_forward(self, ring: torch.Tensor) -> torch.Tensor:
        B, H, W, C = ring.shape
        assert C == 3 and ring.dtype == torch.float32 and ring.is_contiguous()
        out = torch.empty((B, H, W, 8), dtype=torch.float32, device=self.device)
        self.check(self.lib.caelo_respond_forward(self.h, _ptr(ring), B, H, W, _ptr(out), _stream()),
                   "caelo_respond_forward")
        return out

    def select_keypoints(self, ring: torch.Tensor, counter: torch.Tensor, resp: Optional[torch.Tensor],
                         H: int = nLines, W: int = ImgW - CropWidth_SphericalRing, max_kpts: int = nFixedKeyPts):
        """ring [B,rH,rW,C], counter [B,cH,cW] int8|int32, resp [B,H,W,8] or None (fused a1+a2)."""
        B, rH, rW, rC = ring.shape
        assert ring.dtype == torch.float32 and ring.is_contiguous() and counter.is_contiguous()
        kind = {torch.int8: 0, torch.int32: 1}[counter.dtype]
        kpts = torch.empty((B, max_kpts, 3), dtype=torch.float32, device=self.device)
        kpix = torch.empty((B, max_kpts, 2), dtype=torch.int64, device=self.device)
        n = torch.empty((B,), dtype=torch.int32, device=self.device)
        if resp is None:
            rc = self.lib.caelo_respond_select(self.h, _ptr(ring), rC, rH, rW, _ptr(counter), kind,
                                               counter.shape[1], counter.shape[2], H, W, B, max_kpts,
                                               _ptr(kpts), _ptr(kpix), _ptr(n), None, _stream())
        else:
            assert resp.shape == (B, H, W, 8) and resp.is_contiguous()
            rc = self.lib.caelo_select_keypoints(self.h, _ptr(resp), H, W, _ptr(ring), rC, rH, rW,
                                                 _ptr(counter), kind, counter.shape[1], counter.shape[2],
                                                 B, max_kpts, _ptr(kpts), _ptr(kpix), _ptr(n), _stream())
        self.check(rc, "caelo_select_keypoints")
        return kpts, kpix, n

    def project_ring(self, pts: torch.Tensor, pts_offsets: np.ndarray, want=("ring3", "counter_i8"), reuse: bool = False):
        """f1: pts [sumN,4] f32, pts_offsets host int64 [F+1] -> dict of the requested outputs
        (ring5 [F,69,1800,5], counter_i32 [F,69,1800], ring3 [F,64,1792,3], counter_i8) + status [F]."""
        off = np.ascontiguousarray(pts_offsets, np.int64)
        F = off.shape[0] - 1
        assert pts.dtype == torch.float32 and pts.is_contiguous() and pts.shape[1] == 4
        shapes = {"ring5": ((F, ImgH, ImgW, 5), torch.float32), "counter_i32": ((F, ImgH, ImgW), torch.int32),
                  "ring3": ((F, nLines, ImgW - CropWidth_SphericalRing, 3), torch.float32),
                  "counter_i8": ((F, ImgH, ImgW), torch.int8)}
        new = (lambda k, sh, dt: self._scratch("ring_" + k, sh, dt)) if reuse else \
            (lambda k, sh, dt: torch.empty(sh, dtype=dt, device=self.device))
        out = {k: new(k, shapes[k][0], shapes[k][1]) for k in want}
        out["status"] = torch.empty((F,), dtype=torch.int32, device=self.device)
        self.check(self.lib.caelo_project_ring(self.h, _ptr(pts), off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), F,
                                               _ptr(out.get("ring5")), _ptr(out.get("counter_i32")),
                                               _ptr(out.get("ring3")), _ptr(out.get("counter_i8")),
                                               _ptr(out["status"]), _stream()), "caelo_project_ring")
        return out

    def voxelize(self, pts: torch.Tensor, pts_offsets: np.ndarray, cap: Optional[int] = None, want_blocks: bool = False):
        """f2: -> vox int16 [F,3,cap,3], counts int32 [F,4] (+ local0, blocks, cnt when want_blocks), status [F]."""
        off = np.ascontiguousarray(pts_offsets, np.int64)
        F = off.shape[0] - 1
        assert pts.dtype == torch.float32 and pts.is_contiguous() and pts.shape[1] == 4
        if cap is None:
            cap = int(np.diff(off).max())
        out = {"vox": torch.empty((F, 3, cap, 3), dtype=torch.int16, device=self.device),
               "counts": torch.empty((F, 4), dtype=torch.int32, device=self.device),
               "status": torch.empty((F,), dtype=torch.int32, device=self.device)}
        if want_blocks:
            out["local0"] = torch.empty((F, cap, 3), dtype=torch.int16, device=self.device)
            out["blocks"] = torch.empty((F, cap, 3), dtype=torch.int16, device=self.device)
            out["cnt"] = torch.empty((F, cap + 1), dtype=torch.int32, device=self.device)
        self.check(self.lib.caelo_voxelize(self.h, _ptr(pts), off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), F, cap,
                                           _ptr(out["vox"]), _ptr(out["counts"]), _ptr(out.get("local0")),
                                           _ptr(out.get("blocks")), _ptr(out.get("cnt")), _ptr(out["status"]),
                                           _stream()), "caelo_voxelize")
        return out

    def extend_keypoints(self, ring: torch.Tensor, counter: torch.Tensor, kpix: torch.Tensor,
                         n_kpts: Optional[torch.Tensor] = None, zero_counter: bool = False):
        """ExtendKeyPtsInShpericalRing for B frames -> ext [B,K*169,3] f32, n_ext [B] int32."""
        B, rH, rW, rC = ring.shape
        K = kpix.shape[1]
        assert ring.dtype == torch.float32 and ring.is_contiguous() and counter.is_contiguous()
        assert kpix.dtype == torch.int64 and kpix.is_contiguous() and kpix.shape[0] == B
        kind = {torch.int8: 0, torch.int32: 1}[counter.dtype]
        cap = K * 169
        ext = torch.empty((B, cap, 3), dtype=torch.float32, device=self.device)
        n_ext = torch.empty((B,), dtype=torch.int32, device=self.device)
        self.check(self.lib.caelo_extend_keypoints(self.h, _ptr(ring), rC, rH, rW, _ptr(counter), kind, counter.shape[1],
                                                   counter.shape[2], _ptr(kpix), _ptr(n_kpts), B, K, _ptr(ext), cap,
                                                   _ptr(n_ext), 1 if zero_counter else 0, _stream()),
                   "caelo_extend_keypoints")
        return ext, n_ext

    def gather_patches(self, kpts: torch.Tensor, vox: torch.Tensor, vox_offsets: np.ndarray,
                       n_kpts: Optional[torch.Tensor] = None, want_f32: bool = False, want_trunc: bool = False,
                       group: Optional[int] = None, reuse: bool = False):
        """kpts [F,K,3] f32|f64; vox int16 [sumV,3]; vox_offsets host int64 [F*3+1] (rows).
        ``group``: frames per C call (0 / default = the whole batch in one call).  Building and querying the brick
        tables a few frames at a time keeps them in L2 (one frame's tables are ~10 MB, a 33-frame batch's ~340 MB) —
        measured SLOWER: brick_insert 0.18 -> 0.30 ms and gather 0.18 -> 0.23 ms at 8 frames per call; both kernels
        are latency-bound and need the whole batch's parallelism more than the cache."""
        F, K, _ = kpts.shape
        assert kpts.is_contiguous() and vox.dtype == torch.int16 and vox.is_contiguous()
        off = np.ascontiguousarray(vox_offsets, np.int64)
        assert off.shape == (F * 3 + 1,)
        packed = self._scratch("packed", (F, 3, K, 128), torch.int32) if reuse else \
            torch.empty((F, 3, K, 128), dtype=torch.int32, device=self.device)
        f32 = torch.empty((F, 3, K, 16, 16, 16), dtype=torch.float32, device=self.device) if want_f32 else None
        trunc = torch.empty((F, 3, K), dtype=torch.uint8, device=self.device) if want_trunc else None
        if group is None:
            group = int(os.environ.get("CAELO_GATHER_GROUP", "0"))
        group = F if group <= 0 else min(group, F)
        for f0 in range(0, F, group):
            f1 = min(f0 + group, F)
            o = np.ascontiguousarray(off[3 * f0:3 * f1 + 1] - off[3 * f0])
            rc = self.lib.caelo_gather_patches(self.h, _ptr(kpts[f0:f1]), 1 if kpts.dtype == torch.float64 else 0,
                                               _ptr(None if n_kpts is None else n_kpts[f0:f1]), f1 - f0, K,
                                               _ptr(vox[int(off[3 * f0]):]),
                                               o.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                               _ptr(packed[f0:f1]), _ptr(None if f32 is None else f32[f0:f1]),
                                               _ptr(None if trunc is None else trunc[f0:f1]), _stream())
            self.check(rc, "caelo_gather_patches")
        return packed, f32, trunc

    def gather_patches_scans(self, kpts: torch.Tensor, pts: torch.Tensor, pts_offsets: np.ndarray,
                             n_kpts: Optional[torch.Tensor] = None, want_f32: bool = False, want_trunc: bool = False,
                             group: Optional[int] = None, reuse: bool = False):
        """f2+a6 fused: kpts [F,K,3]; pts [sumN,4] f32 raw scans; -> packed, f32, trunc, nvox [F,3], status [F].
        ``group`` as in gather_patches."""
        F, K, _ = kpts.shape
        off = np.ascontiguousarray(pts_offsets, np.int64)
        assert off.shape == (F + 1,) and pts.dtype == torch.float32 and pts.is_contiguous() and kpts.is_contiguous()
        packed = self._scratch("packed", (F, 3, K, 128), torch.int32) if reuse else \
            torch.empty((F, 3, K, 128), dtype=torch.int32, device=self.device)
        f32 = torch.empty((F, 3, K, 16, 16, 16), dtype=torch.float32, device=self.device) if want_f32 else None
        trunc = torch.empty((F, 3, K), dtype=torch.uint8, device=self.device) if want_trunc else None
        nvox = torch.empty((F, 3), dtype=torch.int32, device=self.device)
        status = torch.empty((F,), dtype=torch.int32, device=self.device)
        if group is None:
            group = int(os.environ.get("CAELO_GATHER_GROUP", "0"))
        group = F if group <= 0 else min(group, F)
        for f0 in range(0, F, group):
            f1 = min(f0 + group, F)
            o = np.ascontiguousarray(off[f0:f1 + 1] - off[f0])
            rc = self.lib.caelo_gather_patches_scans(self.h, _ptr(kpts[f0:f1]), 1 if kpts.dtype == torch.float64 else 0,
                                                     _ptr(None if n_kpts is None else n_kpts[f0:f1]), f1 - f0, K,
                                                     _ptr(pts[int(off[f0]):]),
                                                     o.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _ptr(packed[f0:f1]),
                                                     _ptr(None if f32 is None else f32[f0:f1]),
                                                     _ptr(None if trunc is None else trunc[f0:f1]), _ptr(nvox[f0:f1]),
                                                     _ptr(status[f0:f1]), _stream())
            self.check(rc, "caelo_gather_patches_scans")
        return packed, f32, trunc, nvox, status

    # a6 in two steps.  The brick tables are ONE scratch region of the context: a build overwrites the index that
    # ``bricks_gather`` reads.  ``stream`` lets a caller build on another stream, but then the caller has to order the
    # build after the previous batch's gather and before this batch's gather with events — the library does not.
    def bricks_build(self, vox: torch.Tensor, vox_offsets: np.ndarray, stream=None):
        off = np.ascontiguousarray(vox_offsets, np.int64)
        F = (off.shape[0] - 1) // 3
        assert vox.dtype == torch.int16 and vox.is_contiguous() and off.shape == (F * 3 + 1,)
        st = ctypes.c_void_p(stream.cuda_stream) if stream is not None else _stream()
        self.check(self.lib.caelo_bricks_build(self.h, _ptr(vox), off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), F, st),
                   "caelo_bricks_build")
        return F

    def bricks_build_scans(self, pts: torch.Tensor, pts_offsets: np.ndarray, stream=None):
        off = np.ascontiguousarray(pts_offsets, np.int64)
        F = off.shape[0] - 1
        assert pts.dtype == torch.float32 and pts.is_contiguous()
        st = ctypes.c_void_p(stream.cuda_stream) if stream is not None else _stream()
        with torch.cuda.stream(stream) if stream is not None else _nullctx():
            nvox = torch.empty((F, 3), dtype=torch.int32, device=self.device)
            status = torch.empty((F,), dtype=torch.int32, device=self.device)
        self.check(self.lib.caelo_bricks_build_scans(self.h, _ptr(pts), off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), F,
                                                     _ptr(nvox), _ptr(status), st), "caelo_bricks_build_scans")
        return nvox, status

    def bricks_gather(self, kpts: torch.Tensor, n_kpts=None, nvox=None, status=None):
        F, K, _ = kpts.shape
        assert kpts.is_contiguous()
        packed = torch.empty((F, 3, K, 128), dtype=torch.int32, device=self.device)
        self.check(self.lib.caelo_bricks_gather(self.h, _ptr(kpts), 1 if kpts.dtype == torch.float64 else 0, _ptr(n_kpts), F, K,
                                                _ptr(packed), None, None, _ptr(nvox), _ptr(status), _stream()),
                   "caelo_bricks_gather")
        return packed

    def encode_frames(self, packed: torch.Tensor) -> torch.Tensor:
        F, S, K, Wd = packed.shape
        assert S == 3 and Wd == 128 and packed.is_contiguous()
        feat = torch.empty((F, K, 60), dtype=torch.float32, device=self.device)
        self.check(self.lib.caelo_encode_frames(self.h, _ptr(packed), F, K, _ptr(feat), _stream()),
                   "caelo_encode_frames")
        return feat

    def encode_patches(self, patches: torch.Tensor) -> torch.Tensor:
        P = patches.shape[0]
        assert patches.dtype == torch.float32 and patches.is_contiguous() and patches[0].numel() == 4096
        feat = torch.empty((P, 20), dtype=torch.float32, device=self.device)
        status = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self.check(self.lib.caelo_encode_patches(self.h, _ptr(patches), P, _ptr(feat), _ptr(status), _stream()),
                   "caelo_encode_patches")
        self.check(int(status.item()), "caelo_encode_patches(input check)")
        return feat

    def nn_match(self, codes0: torch.Tensor, codes1: torch.Tensor) -> torch.Tensor:
        P, N, D = codes0.shape
        M = codes1.shape[1]
        assert codes1.shape[0] == P and codes1.shape[2] == D
        assert codes0.dtype == torch.float32 and codes0.is_contiguous() and codes1.is_contiguous()
        idx = torch.empty((P, M), dtype=torch.int64, device=self.device)
        self.check(self.lib.caelo_nn_match(self.h, _ptr(codes0), _ptr(codes1), P, N, M, D, _ptr(idx), _stream()),
                   "caelo_nn_match")
        return idx

    def ransac_round(self, pc0, pc1, pair_idx, sample_idx, thr, best_n_in=None, want_counts=False, skip_if_ok=None,
                     out_result=None):
        P, N0, _ = pc0.shape
        N = pc1.shape[1]
        T = sample_idx.shape[1]
        assert sample_idx.is_contiguous() and sample_idx.dtype == torch.int32
        result = out_result if out_result is not None else torch.empty((P, 16), dtype=torch.float32, device=self.device)
        mask = torch.empty((P, N), dtype=torch.uint8, device=self.device)
        counts = torch.empty((P, T), dtype=torch.int32, device=self.device) if want_counts else None
        rc = self.lib.caelo_ransac_round(self.h, _ptr(pc0), N0, _ptr(pc1), N, _ptr(pair_idx), _ptr(sample_idx),
                                         T, _ptr(thr), _ptr(best_n_in), _ptr(skip_if_ok), P, _ptr(result), _ptr(mask),
                                         _ptr(counts), _stream())
        self.check(rc, "caelo_ransac_round")
        return result, mask, counts

    def ransac_ladder(self, pc0, pc1, pair_idx, sample_idx, thresholds):
        """All ladder rounds + the refit for P pairs in one call (no host round trip): sample_idx int32
        [rounds,P,T,4] -> result [P,16], mask [P,N] uint8, rt [P,12], thr_used [P]."""
        P, N0, _ = pc0.shape
        N = pc1.shape[1]
        rounds, _, T, _ = sample_idx.shape
        assert sample_idx.is_contiguous() and sample_idx.dtype == torch.int32 and len(thresholds) == rounds
        result = torch.empty((P, 16), dtype=torch.float32, device=self.device)
        mask = torch.empty((P, N), dtype=torch.uint8, device=self.device)
        rt = torch.empty((P, 12), dtype=torch.float32, device=self.device)
        thr_used = torch.empty((P,), dtype=torch.float32, device=self.device)
        cred = torch.empty((P,), dtype=torch.int32, device=self.device)
        thr = (ctypes.c_float * rounds)(*[float(t) for t in thresholds])
        self.check(self.lib.caelo_ransac_ladder(self.h, _ptr(pc0), N0, _ptr(pc1), N, _ptr(pair_idx), _ptr(sample_idx), T,
                                                rounds, thr, P, _ptr(result), _ptr(mask), _ptr(rt), _ptr(thr_used),
                                                _ptr(cred), _stream()), "caelo_ransac_ladder")
        return result, mask, rt, thr_used

    def draw_samples(self, seeds, n_points: int, rounds: int = 1, rounds_done: int = 0, T: int = MAX_TRIALS):
        """RANSAC sample indices of ``rounds`` ladder rounds for pairs seeded np.random.seed(seeds[i]), generated
        on the device (numpy's legacy MT19937 stream, bit for bit) -> int32 [rounds,P,T,4]."""
        sd = np.ascontiguousarray(seeds, np.int64)
        P = sd.shape[0]
        out = torch.empty((rounds, P, T, 4), dtype=torch.int32, device=self.device)
        self.check(self.lib.caelo_ransac_draw_samples(self.h, sd.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), P,
                                                      int(n_points), T, rounds, rounds_done, _ptr(out), _stream()),
                   "caelo_ransac_draw_samples")
        return out

    def nn3(self, pc0: torch.Tensor, pc1: torch.Tensor, thr: float = 0.0, want_mask: bool = False):
        """Exact 3-D 1-NN of every pc1 row among pc0 (f4, MyICP.py:33-34) -> idx int64 [M], dist float64 [M],
        mask uint8 [M] (dist < thr) and count int32 [1] when ``want_mask``."""
        N, M = pc0.shape[0], pc1.shape[0]
        assert pc0.dtype == torch.float32 and pc1.dtype == torch.float32 and pc0.is_contiguous() and pc1.is_contiguous()
        idx = torch.empty((M,), dtype=torch.int64, device=self.device)
        dist = torch.empty((M,), dtype=torch.float64, device=self.device)
        mask = torch.empty((M,), dtype=torch.uint8, device=self.device) if want_mask else None
        count = torch.empty((1,), dtype=torch.int32, device=self.device) if want_mask else None
        self.check(self.lib.caelo_nn3(self.h, _ptr(pc0), N, _ptr(pc1), M, _ptr(idx), _ptr(dist), float(thr), _ptr(mask),
                                      _ptr(count), _stream()), "caelo_nn3")
        return idx, dist, mask, count

    def transform_points(self, rt: torch.Tensor, pc: torch.Tensor):
        """pc <- R pc + T in place (contract U1); rt dev [12]."""
        assert rt.dtype == torch.float32 and rt.numel() == 12 and pc.dtype == torch.float32 and pc.is_contiguous()
        self.check(self.lib.caelo_transform_points(self.h, _ptr(rt), _ptr(pc), pc.shape[0], _stream()),
                   "caelo_transform_points")

    def icp_batch(self, pc0: torch.Tensor, off0: np.ndarray, pc1: torch.Tensor, off1: np.ndarray, thr0: float, decay: float,
                  small_shift: float, ep: float, max_iter: int, min_iter: int, min_inliers: int = 100):
        """Whole ICPs of B pairs on the device (caelo_icp_batch): pc0 [S0,3] / pc1 [S1,3] f32 = the B target / source
        clouds concatenated (pc1 is updated IN PLACE), off0 / off1 host int64 [B+1].  -> hist [B,max_iter,12] f32,
        hist_n [B,max_iter] int32, state [B,4] float64 (success, iterations, last inlier count, final threshold)."""
        o0, o1 = np.ascontiguousarray(off0, np.int64), np.ascontiguousarray(off1, np.int64)
        B = o0.shape[0] - 1
        assert pc0.dtype == torch.float32 and pc1.dtype == torch.float32 and pc0.is_contiguous() and pc1.is_contiguous()
        assert o1.shape[0] == B + 1 and int(o0[-1]) == pc0.shape[0] and int(o1[-1]) == pc1.shape[0]
        hist = torch.empty((B, max_iter, 12), dtype=torch.float32, device=self.device)
        hist_n = torch.empty((B, max_iter), dtype=torch.int32, device=self.device)
        state = torch.empty((B, 4), dtype=torch.float64, device=self.device)
        self.check(self.lib.caelo_icp_batch(self.h, _ptr(pc0), o0.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _ptr(pc1),
                                            o1.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), B, float(thr0), float(decay),
                                            float(small_shift), float(ep), int(max_iter), int(min_iter), int(min_inliers),
                                            _ptr(hist), _ptr(hist_n), _ptr(state), _stream()), "caelo_icp_batch")
        return hist, hist_n, state

    def kabsch(self, pc0, pc1, pair_idx=None, mask=None, skip_if_ok=None, out_rt=None):
        P, N0, _ = pc0.shape
        N = pc1.shape[1]
        rt = out_rt if out_rt is not None else torch.empty((P, 12), dtype=torch.float32, device=self.device)
        cred = torch.empty((P,), dtype=torch.int32, device=self.device)
        self.check(self.lib.caelo_kabsch(self.h, _ptr(pc0), N0, _ptr(pc1), N, _ptr(pair_idx), _ptr(mask),
                                         _ptr(skip_if_ok), P, _ptr(rt), _ptr(cred), _stream()), "caelo_kabsch")
        return rt, cred


def _default_weights(which: str):
    z = np.load(os.path.join(WEIGHT_DIR, which + ".npz"))
    return {k: z[k] for k in z.files}


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.device_count() > 1 else 0
        _default_ctx = Context(dev)
    return _default_ctx


def _dev(a: np.ndarray, dtype=None) -> torch.Tensor:
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype))
    return torch.from_numpy(a).to(default_context().device, non_blocking=False)


# ------------------------------------------------------------------------------------------
# Keras stand-in
# ------------------------------------------------------------------------------------------
class B200Model:
    """What ``keras.models.load_model`` returns here: ``predict(ndarray) -> ndarray``."""

    def __init__(self, kind: str, weights: dict, ctx: Optional[Context] = None):
        if kind not in ("respond", "encoder"):
            raise ValueError(kind)
        self.kind = kind
        self.ctx = ctx or default_context()
        self.weights = weights
        self._bind()

    def _bind(self):
        """The weights are state of the (shared) context: a model loaded later replaces them.  Every model remembers its
        own and puts them back before it predicts, so two models of one kind can coexist (the setter waits for queued
        work that still reads the old ones)."""
        owners = self.ctx.__dict__.setdefault("_weight_owner", {})
        if owners.get(self.kind) is not self:
            (self.ctx.set_respond_weights if self.kind == "respond" else self.ctx.set_encoder_weights)(self.weights)
            owners[self.kind] = self

    def predict(self, x, batch_size=None, verbose=0):
        self._bind()
        x = np.asarray(x, dtype=np.float32)
        if self.kind == "respond":
            if x.ndim != 4 or x.shape[3] != 3:
                raise ValueError("expected input (B,H,W,3), got %r" % (x.shape,))
            out = self.ctx.respond_forward(_dev(x))
            return out.cpu().numpy()
        if x.ndim == 5 and x.shape[1:] == (16, 16, 16, 1):
            x = x.reshape(x.shape[0], 16, 16, 16)
        if x.ndim != 4 or x.shape[1:] != (16, 16, 16):
            raise ValueError("expected input (K,16,16,16,1), got %r" % (x.shape,))
        if x.shape[0] == 0:
            return np.zeros((0, 20), np.float32)
        return self.ctx.encode_patches(_dev(x)).cpu().numpy()


def load_model(path: str, ctx: Optional[Context] = None) -> B200Model:
    """Drop-in for ``keras.models.load_model`` on the two inference networks CAE-LO ships
    (TrainedModels/SphericalRingPCRespondLayer.h5, EncoderModel4VoxelPatch.h5)."""
    if path.endswith(".npz"):
        z = np.load(path)
        w = {k: z[k] for k in z.files}
        layers = None
    else:
        w, _sha = read_keras_weights(path)
        layers = layer_summary(read_model_config(path))
    if "conv2d_1/kernel:0" in w and "conv3d_1/kernel:0" not in w:
        if layers is not None:
            acts = [l[2] for l in layers if l[0] == "Conv2D"]
            if acts != ["relu", "relu"]:
                raise _lib.CaeloError("unsupported respond-layer activations %r" % acts)
        return B200Model("respond", w, ctx)
    if "conv3d_1/kernel:0" in w and "dense_2/kernel:0" in w:
        if layers is not None:
            acts = [l[2] for l in layers if l[0] in ("Conv3D", "Dense")]
            if acts != ["tanh"] * 5:
                raise _lib.CaeloError("unsupported encoder activations %r (the kernel implements the "
                                      "shipped all-tanh EncoderModel4VoxelPatch.h5)" % acts)
        return B200Model("encoder", w, ctx)
    raise _lib.CaeloError("%s is neither the respond layer nor the voxel-patch encoder" % path)


# ------------------------------------------------------------------------------------------
# per-scan pre-stages (f1, f2)
# ------------------------------------------------------------------------------------------
def ProjectPC2SphericalRing(PC):
    """SphericalRing.py:72 — (Image_float (69,1800,5) f32, GridCounter (69,1800) int32)."""
    PC = np.asarray(PC)
    assert PC.shape[0] > 3 and PC.shape[1] == 4
    ctx = default_context()
    pts = _dev(PC, np.float32)
    out = ctx.project_ring(pts, np.array([0, pts.shape[0]], np.int64), want=("ring5", "counter_i32"))
    if int(out["status"].item()):
        raise IndexError("index %d is out of bounds for axis 1 with size %d" % (ImgW, ImgW))
    return out["ring5"][0].cpu().numpy(), out["counter_i32"][0].cpu().numpy()


def Voxelization(PC):
    """Voxel.py:100 — (Blocks, VoxelModel1, VoxelModel2, avlBlocksList, cntVoxelsLength, AllVoxels,
    AllVoxels0, AllVoxels1, AllVoxels2).  The three dense/nested containers (Blocks: 560k python lists,
    VoxelModel1: a 286 MB int8 grid) are never read on the odometry path and come back as None; the six
    arrays BatchVoxelization.py:61 stores are exact, order included."""
    PC = np.asarray(PC)
    ctx = default_context()
    pts = np.zeros((PC.shape[0], 4), np.float32)
    pts[:, :3] = PC[:, :3]
    d = ctx.voxelize(_dev(pts), np.array([0, pts.shape[0]], np.int64), want_blocks=True)
    if int(d["status"].item()):
        raise IndexError("a point indexes outside the block grid")
    n0, n1, n2, nb = (int(c) for c in d["counts"][0].cpu().numpy())
    vox = d["vox"][0]
    return (None, None, None, d["blocks"][0, :nb].cpu().numpy(), d["cnt"][0, :nb + 1].cpu().numpy(),
            d["local0"][0, :n0].cpu().numpy(), vox[0, :n0].cpu().numpy(), vox[1, :n1].cpu().numpy(),
            vox[2, :n2].cpu().numpy())


# ------------------------------------------------------------------------------------------
# keypoints
# ------------------------------------------------------------------------------------------
def _select(SphericalRing, GridCounter, RespondImg):
    ctx = default_context()
    ring = np.asarray(SphericalRing, dtype=np.float32)
    cnt = np.asarray(GridCounter)
    if cnt.dtype != np.int8:
        cnt = cnt.astype(np.int32, copy=False)
    resp = None
    if RespondImg is not None:
        r = np.asarray(RespondImg, dtype=np.float32)
        H, W = r.shape[0], r.shape[1]
        resp = _dev(r[None])
    else:
        H, W = nLines, ImgW - CropWidth_SphericalRing
    kpts, kpix, n = ctx.select_keypoints(_dev(ring[None]), _dev(cnt[None]), resp, H, W, nFixedKeyPts)
    n = int(n.item())
    KeyPts = kpts[0, :n].cpu().numpy()
    KeyPixels = kpix[0, :n].cpu().numpy()
    PlanarPts = np.array([], dtype=np.float32)
    assert KeyPts.shape[0] > 50  # SphericalRing.py:286
    return KeyPts, KeyPixels, PlanarPts


def GetKeyPtsByAE(SphericalRing, GridCounter, RespondImg):
    """SphericalRing.py:113 — (KeyPts (n,3) f32 ascending by score, KeyPixels (n,2) int64, PlanarPts)."""
    return _select(SphericalRing, GridCounter, RespondImg)


def ExtendKeyPtsInShpericalRing(SphericalRing, GridCounter, KeyPixels):
    """SphericalRing.py:294 — ExtendedKeyPts (n,3) f32; like the reference it zeroes the windows of the
    caller's GridCounter in place."""
    ctx = default_context()
    ring = np.asarray(SphericalRing, dtype=np.float32)
    px = np.ascontiguousarray(np.asarray(KeyPixels).reshape(-1, 2), np.int64)
    if px.shape[0] == 0:
        return np.zeros((0, 3), np.float32)
    cnt = GridCounter if GridCounter.dtype == np.int8 else GridCounter.astype(np.int32, copy=False)
    d_cnt = _dev(cnt[None])
    ext, n = ctx.extend_keypoints(_dev(ring[None]), d_cnt, _dev(px[None]), None, zero_counter=True)
    GridCounter[...] = d_cnt[0].cpu().numpy()
    return ext[0, :int(n.item())].cpu().numpy()


def GetKeyPtsFromRing(SphericalRing, GridCounter):
    """Fused a1+a2: what GetKeyPtsFromRawFileName computes once the .mat is loaded."""
    return _select(SphericalRing, GridCounter, None)


def GetKeyPtsFromRawFileName(rawFileFullPath, RespondLayer=None):
    """SphericalRing.py:389 — loads <seq>/SphericalRing/<name>.mat next to the raw file."""
    from scipy import io
    baseDir = os.path.dirname(os.path.dirname(rawFileFullPath))
    mat = io.loadmat(os.path.join(baseDir, "SphericalRing", os.path.basename(rawFileFullPath) + ".mat"))
    return GetKeyPtsFromRing(mat["SphericalRing"], mat["GridCounter"])


# ------------------------------------------------------------------------------------------
# patches + descriptors
# ------------------------------------------------------------------------------------------
def _vox_cat(AllVoxels0, AllVoxels1, AllVoxels2):
    lists = [np.ascontiguousarray(np.asarray(v), np.int16).reshape(-1, 3) for v in (AllVoxels0, AllVoxels1, AllVoxels2)]
    off = np.zeros(4, np.int64)
    off[1:] = np.cumsum([l.shape[0] for l in lists])
    if min(l.shape[0] for l in lists) < 496:
        raise ValueError("Expected n_neighbors <= n_samples_fit")  # what sklearn raises (Voxel.py:195)
    return np.concatenate(lists, 0), off


def _gather(Pts, AllVoxels0, AllVoxels1, AllVoxels2, want_f32):
    ctx = default_context()
    P = np.asarray(Pts)
    if P.dtype != np.float64:
        P = P.astype(np.float32, copy=False)
    vox, off = _vox_cat(AllVoxels0, AllVoxels1, AllVoxels2)
    return ctx.gather_patches(_dev(P[None]), _dev(vox), off, None, want_f32=want_f32)


def GetPatchesList(Pts, AllVoxels0, AllVoxels1, AllVoxels2):
    """Voxel.py:177 — (Pts, [3 x (K,16,16,16,1) float32])."""
    _packed, f32, _ = _gather(Pts, AllVoxels0, AllVoxels1, AllVoxels2, True)
    K = np.asarray(Pts).shape[0]
    out = f32[0].cpu().numpy().reshape(3, K, 16, 16, 16, 1)
    return Pts, [out[0], out[1], out[2]]


def GetFeaturesFromPatches(PatchEncoder, PatchesList):
    """Match.py:130 — np.c_[predict(p0), predict(p1), predict(p2)]."""
    return np.c_[PatchEncoder.predict(PatchesList[0]), PatchEncoder.predict(PatchesList[1]),
                 PatchEncoder.predict(PatchesList[2])]


def GetFeaturesAtKeyPts(Pts, AllVoxels0, AllVoxels1, AllVoxels2):
    """a6+a3 without materialising float32 patches: GetFeaturesFromPatches(GetPatchesList(...))."""
    ctx = default_context()
    packed, _, _ = _gather(Pts, AllVoxels0, AllVoxels1, AllVoxels2, False)
    return ctx.encode_frames(packed)[0].cpu().numpy()


# ------------------------------------------------------------------------------------------
# match + pose
# ------------------------------------------------------------------------------------------
def SolveRT(Pairs0, Pairs1):
    """Match.py:138 — (R (3,3) f32, T (3,1) f32, isCredible)."""
    ctx = default_context()
    p0 = _dev(np.asarray(Pairs0, np.float32)[None])
    p1 = _dev(np.asarray(Pairs1, np.float32)[None])
    rt, cred = ctx.kabsch(p0, p1)
    rt = rt.cpu().numpy()[0]
    return rt[:9].reshape(3, 3).copy(), rt[9:].reshape(3, 1).copy(), int(cred.item())


def _ransac(ctx, pc0, pc1, pair_idx, N, verbose=False):
    """RANSAC4RT's threshold ladder around caelo_ransac_round; consumes np.random exactly like
    the reference (Match.py:181-214): 4 doubles per trial actually run."""
    thr = 0.4
    best_n = 0
    R_star = np.eye(3, dtype=np.float64)
    T_star = np.zeros((3, 1), dtype=np.float64)
    mask_star = None
    ok = False
    used = 0
    while True:
        state = np.random.get_state()
        u = np.random.random((MAX_TRIALS, 4))
        idx = np.array(u * N, dtype=np.int32)
        res, mask, _ = ctx.ransac_round(pc0, pc1, pair_idx, _dev(idx[None]),
                                        _dev(np.array([thr], np.float32)),
                                        _dev(np.array([best_n], np.int32)))
        r = res.cpu().numpy()[0]
        used = int(r[13])
        np.random.set_state(state)
        np.random.random((used * 4,))
        if int(r[15]) >= 0:
            best_n = int(r[14])
            R_star = r[:9].reshape(3, 3).copy()
            T_star = r[9:12].reshape(3, 1).copy()
            mask_star = mask
        if r[12] != 0:
            ok = True
            break
        thr = 2 * thr
        if thr > 2.0:
            if verbose:
                print('failed when residual =', thr)
            thr = thr / 2
            break
    if verbose:
        print('cntItersRANSAC =', used)
        print('residualThreshold =', thr)
        print('nInliers/nFilteredKeyPts1 =', best_n, '/', N, '=', round(best_n / N, 3))
    return R_star, T_star, ok, mask_star, thr, used


def RANSAC4RT(Pairs0, Pairs1, Weights0=None, Weights1=None):
    """Match.py:162 — (R*, T*, isSuccess, inlierMask bool (N,), residualThreshold)."""
    ctx = default_context()
    P0 = np.asarray(Pairs0, np.float32)
    N = P0.shape[0]
    R, T, ok, mask, thr, _used = _ransac(ctx, _dev(P0[None]), _dev(np.asarray(Pairs1, np.float32)[None]), None, N)
    m = np.zeros((N,), dtype=bool) if mask is None else mask[0].cpu().numpy().astype(bool)
    return R, T, ok, m, thr


def SolveRelativePose(OriPC0, OriCodes0, Weights0, OriPC1, OriCodes1, Weights1):
    """Match.py:241 — (R, T, isSuccess, inliersIdx0, inliersIdx1, residualThreshold);
    x0 ~= R x1 + T.  Weights are ignored, as in the reference (overwritten with ones, :265-266)."""
    return solve_relative_pose(default_context(), OriPC0, OriCodes0, OriPC1, OriCodes1)[:6]


def solve_relative_pose(ctx, OriPC0, OriCodes0, OriPC1, OriCodes1):
    """SolveRelativePose on a given context; also returns the number of RANSAC trials of the last ladder round."""
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32)[None])).to(ctx.device)
    pc0, pc1, c0, c1 = dev(OriPC0), dev(OriPC1), dev(OriCodes0), dev(OriCodes1)
    N = pc1.shape[1]
    pair_idx = ctx.nn_match(c0, c1)
    R, T, ok, mask, thr, used = _ransac(ctx, pc0, pc1, pair_idx, N)
    if mask is None:
        e = np.zeros((0,), np.int64)
        return R, T, ok, e, e.copy(), thr, used
    m = mask[0].cpu().numpy().astype(bool)
    pidx = pair_idx[0].cpu().numpy()
    inliersIdx0 = pidx[m]
    inliersIdx1 = np.arange(N)[m]
    if inliersIdx0.shape[0] == 0:
        return R, T, ok, inliersIdx0, inliersIdx1, thr, used
    rt, _ = ctx.kabsch(pc0, pc1, pair_idx, mask)
    rt = rt.cpu().numpy()[0]
    return rt[:9].reshape(3, 3).copy(), rt[9:].reshape(3, 1).copy(), ok, inliersIdx0, inliersIdx1, thr, used


# ------------------------------------------------------------------------------------------
# f4: ICP on the extended key points (MyICP.py)
# ------------------------------------------------------------------------------------------
RADIAN2DEGREE = 180.0 / np.pi


def RotateMat2EulerAngle_XYZ(R):
    """Transformations.py:181-186 (degrees)."""
    import math
    angles = np.zeros((3,))
    angles[0] = math.atan2(R[2, 1], R[2, 2]) * RADIAN2DEGREE
    angles[1] = math.atan2(-R[2, 0], math.sqrt(math.pow(R[2, 1], 2) + math.pow(R[2, 2], 2))) * RADIAN2DEGREE
    angles[2] = math.atan2(R[1, 0], R[0, 0]) * RADIAN2DEGREE
    return angles


def GetPtsInliners(PC0, PC1, inlierThreshold):
    """MyICP.py:76-85 — nearest PC0 point of every PC1 point, pairs closer than the threshold."""
    ctx = default_context()
    p0, p1 = np.ascontiguousarray(PC0, np.float32), np.ascontiguousarray(PC1, np.float32)
    idx, _dist, mask, _ = ctx.nn3(_dev(p0), _dev(p1), inlierThreshold, want_mask=True)
    idx1 = mask.cpu().numpy().astype(bool)
    idx0 = idx.cpu().numpy()[idx1]
    return PC0[idx0, :], PC1[idx1, :]


def _icp_replay(hist, hist_n, maxIterTimes, minIterTimes, inlierThreshold, smallShiftThreshold, decay_rate, ep, min_inliers=100):
    """The reference's loop (MyICP.py:31-70) over the [R|T] and inlier count the device recorded for every iteration:
    R*, T* accumulated with the reference's own numpy expressions (float32 R, T into float64 R*, T*), and every
    decision — failure, convergence, threshold decay — taken again on the host.
    -> (R_star, T_star, isSuccess, iterations, last inlier count, final threshold)."""
    R_star = np.eye(3, dtype=np.float64)
    T_star = np.zeros((3, 1), dtype=np.float64)
    n_in, it_done = 0, 0
    for iIter in range(maxIterTimes):
        n_in, it_done = int(hist_n[iIter]), iIter + 1
        if n_in < min_inliers:
            return R_star, T_star, False, it_done, n_in, inlierThreshold
        R, T = hist[iIter, :9].reshape(3, 3), hist[iIter, 9:12].reshape(3, 1)
        R_star = np.dot(R, R_star)
        T_star = np.dot(R, T_star) + T
        normEulers = np.linalg.norm(RotateMat2EulerAngle_XYZ(R))
        normT = np.linalg.norm(T)
        if iIter >= minIterTimes and normEulers < ep and normT < ep:
            break
        if normEulers < smallShiftThreshold and normT < smallShiftThreshold:
            inlierThreshold *= decay_rate
    return R_star, T_star, True, it_done, n_in, inlierThreshold


def icp_batch(PC0s, PC1s, maxIterTimes=50, minIterTimes=20 - 1, inlierThreshold=0.5, smallShiftThreshold=0.05,
              decay_rate=0.9, ep=0.001, ctx: Optional[Context] = None):
    """MyICP.ICP (MyICP.py:28-73) for a whole batch of cloud pairs at once: ONE call runs every pair's iterations on the
    device (nearest neighbours through a grid index, SolveRT, point update, loop control) and ONE copy brings the
    per-iteration [R|T] back; the host then accumulates R*, T* as the reference does and re-takes every loop decision
    from the recorded data.  Should a decision differ from the device's (its Euler angles come from CUDA's atan2, the
    host's from libm — they can only disagree when a norm sits within an ulp of a threshold) that pair is redone
    with the host-driven iteration of ``_icp_stepwise``.  -> [(R_star (3,3) f64, T_star (3,1) f64, isSuccess, info)]."""
    ctx = ctx or default_context()
    p0 = [np.ascontiguousarray(p, np.float32).reshape(-1, 3) for p in PC0s]
    p1 = [np.ascontiguousarray(p, np.float32).reshape(-1, 3) for p in PC1s]
    assert len(p0) == len(p1) and len(p0) > 0 and all(a.shape[0] > 0 and b.shape[0] > 0 for a, b in zip(p0, p1))
    off0, off1 = np.zeros(len(p0) + 1, np.int64), np.zeros(len(p0) + 1, np.int64)
    off0[1:], off1[1:] = np.cumsum([a.shape[0] for a in p0]), np.cumsum([a.shape[0] for a in p1])
    d0 = torch.from_numpy(np.concatenate(p0, 0)).to(ctx.device)
    d1 = torch.from_numpy(np.concatenate(p1, 0)).to(ctx.device)
    hist, hist_n, state = ctx.icp_batch(d0, off0, d1, off1, inlierThreshold, decay_rate, smallShiftThreshold, ep,
                                        maxIterTimes, minIterTimes)
    hist, hist_n, state = hist.cpu().numpy(), hist_n.cpu().numpy(), state.cpu().numpy()      # the batch's one sync
    out = []
    for b in range(len(p0)):
        R, T, ok, iters, n_in, thr = _icp_replay(hist[b], hist_n[b], maxIterTimes, minIterTimes, inlierThreshold,
                                                 smallShiftThreshold, decay_rate, ep)
        if (ok, iters, n_in, thr) != (bool(state[b, 0]), int(state[b, 1]), int(state[b, 2]), float(state[b, 3])):
            info = {}
            R, T, ok = _icp_stepwise(ctx, p0[b], p1[b], maxIterTimes, minIterTimes, inlierThreshold, smallShiftThreshold,
                                     decay_rate, ep, info)
            info["redone_on_host"] = True
        else:
            info = dict(iters=iters, inliers=n_in, threshold=thr)
        out.append((R, T, ok, info))
    return out


def ICP(PC0, PC1, maxIterTimes=50, minIterTimes=20 - 1, inlierThreshold=0.5, smallShiftThreshold=0.05, decay_rate=0.9,
        ep=0.001, info=None):
    """MyICP.py:28-73, same arguments and return values (R_star (3,3) f64, T_star (3,1) f64, isSuccess) and the same
    progress line on stdout; a batch of one through ``icp_batch``."""
    R_star, T_star, ok, inf = icp_batch([PC0], [PC1], maxIterTimes, minIterTimes, inlierThreshold, smallShiftThreshold,
                                        decay_rate, ep)[0]
    print('ICP iters:', inf["iters"], ',  inliers:', inf["inliers"], ',  inlierThreshold:', round(inf["threshold"], 5))
    if info is not None:
        info.update(iters=inf["iters"], inliers=inf["inliers"], threshold=inf["threshold"])
    return R_star, T_star, ok


def _icp_stepwise(ctx, PC0, PC1, maxIterTimes, minIterTimes, inlierThreshold, smallShiftThreshold, decay_rate, ep, info=None):
    """One ICP with the loop on the host (one small D2H per iteration): brute-force exact 1-NN (caelo_nn3), SolveRT
    (caelo_kabsch), point update (caelo_transform_points).  The fallback of ``icp_batch`` and its cross-check in the
    tests."""
    pc0 = torch.from_numpy(np.ascontiguousarray(PC0, np.float32)).to(ctx.device)
    pc1 = torch.from_numpy(np.ascontiguousarray(PC1, np.float32)).to(ctx.device).clone()
    hist = np.zeros((maxIterTimes, 12), np.float32)
    hist_n = np.zeros((maxIterTimes,), np.int32)
    thr = inlierThreshold
    for iIter in range(maxIterTimes):
        idx, _dist, mask, count = ctx.nn3(pc0, pc1, thr, want_mask=True)
        rt, _cred = ctx.kabsch(pc0[None], pc1[None], idx[None], mask[None])
        host = torch.cat([rt[0], count.to(torch.float32)]).cpu().numpy()
        hist_n[iIter] = int(host[12])
        if hist_n[iIter] < 100:
            break
        hist[iIter] = host[:12]
        ctx.transform_points(rt[0], pc1)
        # the decisions that steer the next search (same expressions as _icp_replay)
        R, T = hist[iIter, :9].reshape(3, 3), hist[iIter, 9:12].reshape(3, 1)
        normEulers, normT = np.linalg.norm(RotateMat2EulerAngle_XYZ(R)), np.linalg.norm(T)
        if iIter >= minIterTimes and normEulers < ep and normT < ep:
            break
        if normEulers < smallShiftThreshold and normT < smallShiftThreshold:
            thr *= decay_rate
    R_star, T_star, ok, iters, n_in, thr = _icp_replay(hist, hist_n, maxIterTimes, minIterTimes, inlierThreshold,
                                                       smallShiftThreshold, decay_rate, ep)
    if info is not None:
        info.update(iters=iters, inliers=n_in, threshold=thr)
    return R_star, T_star, ok


def _nn3_inliers(ctx, pc0_dev, pc1_host, thr):
    """(idx0 of the inlier pairs, bool mask over pc1) for dist < thr (MyICP.py:77-82)."""
    idx, _dist, mask, _ = ctx.nn3(pc0_dev, _dev(np.ascontiguousarray(pc1_host, np.float32)), thr, want_mask=True)
    m = mask.cpu().numpy().astype(bool)
    return idx.cpu().numpy()[m], m


def GetPlanarPtsInliners(PtsWithNorm0, PtsWithNorm1, inlierThreshold0, inlierThreshold1):
    """MyICP.py:88-113 — nearest frame-0 planar point of every frame-1 planar point (device), then the foot point
    of the frame-0 point on the frame-1 tangent plane and the point-to-plane gate, evaluated with the reference's
    own numpy expressions on the host (at most 2000 rows)."""
    ctx = default_context()
    PC0, PC1, Norms1 = PtsWithNorm0[:, 0:3], PtsWithNorm1[:, 0:3], PtsWithNorm1[:, 3:6]
    idx0, idx1 = _nn3_inliers(ctx, _dev(np.ascontiguousarray(PC0, np.float32)), PC1, inlierThreshold1)
    inliers0 = PC0[idx0, :]
    inliers1 = PC1[idx1, :]
    norms1 = Norms1[idx1, :]
    vetors = inliers0 - inliers1
    dist2Planes = np.sum(norms1 * vetors, axis=1)
    pedals = inliers1 + norms1 * np.tile(dist2Planes.reshape(dist2Planes.shape[0], 1), [1, 3])
    distances = np.linalg.norm((pedals - inliers1), axis=1)
    idx = (distances < inlierThreshold0).flatten()
    return pedals[idx, :], inliers1[idx, :]


def ICP_Pt2PtAndPt2Plane(PC0, PC1, PtsWithNorm0, PtsWithNorm1, maxIterTimes=50, minIterTimes=20 - 1,
                         inlierThreshold0=0.5, decay_rate0=0.9, inlierThreshold1=2.0, decay_rate1=0.5,
                         smallShiftThreshold=0.1, ep=0.01, info=None):
    """MyICP.py:127-201, same arguments, return values, progress line, use of the global np.random stream (planar
    points beyond 2000 are subsampled) and in-place update of PtsWithNorm1's coordinates.  The two nearest-neighbour
    searches, SolveRT on the stacked pairs and the point updates run on the device.  With EMPTY planar arrays — what
    the shipped pipeline produces — the reference raises (IndexError on the 1-D array); so does this."""
    ctx = default_context()
    R_star = np.eye(3, dtype=np.float64)
    T_star = np.zeros((3, 1), dtype=np.float64)
    nMaxPts = 2000
    if PtsWithNorm1.shape[0] > nMaxPts:
        RandIdxes = np.random.random((nMaxPts,))
        RandIdxes = RandIdxes * (PtsWithNorm1.shape[0])
        RandIdxes = np.array(RandIdxes, dtype=np.int32)
        PtsWithNorm1 = PtsWithNorm1[RandIdxes, :]
    PtsWithNorm0[:, 0:3], PtsWithNorm1[:, 0:3]           # raises like the reference on the empty 1-D arrays
    pc0_host = np.ascontiguousarray(PC0, np.float32)
    pc0 = _dev(pc0_host)
    pc1 = _dev(np.ascontiguousarray(PC1, np.float32)).clone()
    isSuccess = True
    minNumOfInputPts = 200
    n_pts = n_pl = 0
    iIter = -1
    for iIter in range(maxIterTimes):
        idx, _dist, mask, _ = ctx.nn3(pc0, pc1, inlierThreshold0, want_mask=True)
        m = mask.cpu().numpy().astype(bool)
        pc1_host = pc1.cpu().numpy()
        inliers0_pts, inliers1_pts = pc0_host[idx.cpu().numpy()[m], :], pc1_host[m, :]
        inliers0_planarPts, inliers1_planarPts = GetPlanarPtsInliners(PtsWithNorm0, PtsWithNorm1, inlierThreshold0,
                                                                      inlierThreshold1)
        n_pts, n_pl = inliers0_pts.shape[0], inliers0_planarPts.shape[0]
        inliers0 = np.r_[inliers0_pts, inliers0_planarPts]
        inliers1 = np.r_[inliers1_pts, inliers1_planarPts]
        if inliers0.shape[0] < minNumOfInputPts:
            if iIter < 1:
                isSuccess = False
            break
        R, T, _isCredible = SolveRT(inliers0, inliers1)
        rt = _dev(np.r_[R.ravel(), T.ravel()].astype(np.float32))
        ctx.transform_points(rt, pc1)
        pl = _dev(np.ascontiguousarray(PtsWithNorm1[:, 0:3], np.float32))
        ctx.transform_points(rt, pl)
        PtsWithNorm1[:, 0:3] = pl.cpu().numpy()
        R_star = np.dot(R, R_star)
        T_star = np.dot(R, T_star) + T
        eulers = RotateMat2EulerAngle_XYZ(R)
        normEulers = np.linalg.norm(eulers)
        normT = np.linalg.norm(T)
        if iIter >= minIterTimes:
            if normEulers < ep and normT < ep:
                break
        if normEulers < smallShiftThreshold and normT < smallShiftThreshold:
            inlierThreshold0 *= decay_rate0
            inlierThreshold1 *= decay_rate1
    print('ICP iters:', iIter + 1, ', inliers0:', n_pts, ', inliers1:', n_pl,
          ', th0:', round(inlierThreshold0, 5), ', th1:', round(inlierThreshold1, 5))
    if info is not None:
        info.update(iters=iIter + 1, inliers0=n_pts, inliers1=n_pl, th0=inlierThreshold0, th1=inlierThreshold1)
    return R_star, T_star, isSuccess
