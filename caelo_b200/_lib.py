"""ctypes binding of libcaelo_b200.so (include/caelo.h).  No fallback: if the shared library
is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CAELO_SO_PATH") or os.path.join(_HERE, "libcaelo_b200.so")   # CAELO_SO_PATH: an experimental build (A/B timing)


class CaeloError(RuntimeError):
    pass


# name -> (restype, argtypes); one entry per symbol declared in include/caelo.h
SIGNATURES = {
    "caelo_version": (c_int, []),
    "caelo_error_string": (c_char_p, [c_int]),
    "caelo_last_cuda_error": (c_char_p, [c_void_p]),
    "caelo_create": (c_int, [c_int, POINTER(c_void_p)]),
    "caelo_destroy": (c_int, [c_void_p]),
    "caelo_num_sms": (c_int, [c_void_p]),
    "caelo_launch_count": (c_int64, [c_void_p]),
    "caelo_profile_enable": (c_int, [c_void_p, c_int]),
    "caelo_profile_fetch": (c_int, [c_void_p, c_char_p, c_int]),
    "caelo_set_respond_weights": (c_int, [c_void_p] + [c_void_p] * 4),
    "caelo_set_encoder_weights": (c_int, [c_void_p] + [c_void_p] * 10),
    "caelo_respond_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "caelo_select_keypoints": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                       c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p]),
    "caelo_respond_select": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "caelo_gather_patches": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                     POINTER(c_int64), c_void_p, c_void_p, c_void_p, c_void_p]),
    "caelo_encode_patches": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "caelo_encode_packed": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "caelo_encode_frames": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "caelo_nn_match": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "caelo_ransac_round": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                   c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "caelo_ransac_ladder": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                    POINTER(c_float), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "caelo_ransac_draw_samples": (c_int, [c_void_p, POINTER(c_int64), c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "caelo_project_ring": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "caelo_voxelize": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p]),
    "caelo_gather_patches_scans": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                           POINTER(c_int64), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p]),
    "caelo_bricks_build": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int, c_void_p]),
    "caelo_bricks_build_scans": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int, c_void_p, c_void_p, c_void_p]),
    "caelo_bricks_gather": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "caelo_extend_keypoints": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "caelo_nn3": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, ctypes.c_double, c_void_p, c_void_p,
                          c_void_p]),
    "caelo_transform_points": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "caelo_icp_batch": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_void_p, POINTER(c_int64), c_int] +
                        [ctypes.c_double] * 4 + [c_int] * 3 + [c_void_p] * 4),
    "caelo_debug_set_timeline": (c_int, [c_void_p, c_void_p]),
    "caelo_debug_nn_last": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "caelo_debug_set_nn_margin": (c_int, [c_void_p, c_float]),
    "caelo_debug_umma": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_void_p, c_void_p]),
    "caelo_kabsch": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                             c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load():
    """dlopen the in-tree shared library and declare every prototype (no compute call)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(SO_PATH):
        raise CaeloError("%s is missing: build it with `python -m caelo_b200.build` "
                         "(there is no CPU fallback)" % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, ctx=None, what: str = ""):
    if rc == 0:
        return
    lib = load()
    msg = lib.caelo_error_string(rc).decode()
    if rc == -1 and ctx is not None:
        msg += ": " + lib.caelo_last_cuda_error(ctx).decode()
    raise CaeloError("%s failed (%d): %s" % (what or "caelo call", rc, msg))
