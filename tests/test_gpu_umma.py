"""GPU: pin the tcgen05 facts the encoder kernels rely on — K-major no-swizzle smem descriptor
semantics (SBO = byte stride between 8-row groups, LBO = byte stride between the two 8-element
K chunks) and the TMEM accumulator layout for M=128 (row m -> lane m) and M=64 (row m -> lane
32*(m/16) + m%16, second tile interleaved at lane offset 16)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from caelo_b200 import api
    return api.default_context()


def _run(ctx, M, N, K, sbo_a, lbo_a, sbo_b, lbo_b, lane_off=0, seed=0):
    import torch
    rng = np.random.default_rng(seed)
    A = rng.integers(-4, 5, (M, K)).astype(np.float16)
    B = rng.integers(-4, 5, (N, K)).astype(np.float16)
    dA = torch.from_numpy(A).cuda()
    dB = torch.from_numpy(B).cuda()
    dump = torch.zeros((128, 64), dtype=torch.float32, device="cuda")
    rc = ctx.lib.caelo_debug_umma(ctx.h, ctypes.c_void_p(dA.data_ptr()), ctypes.c_void_p(dB.data_ptr()), M, N, K,
                                  sbo_a, lbo_a, sbo_b, lbo_b, lane_off, ctypes.c_void_p(dump.data_ptr()), None)
    ctx.check(rc, "caelo_debug_umma")
    torch.cuda.synchronize()
    return dump.cpu().numpy(), A.astype(np.float32) @ B.astype(np.float32).T


def _m64_lanes(lane_off=0):
    m = np.arange(64)
    return 32 * (m // 16) + m % 16 + lane_off


@pytest.mark.parametrize("N,K", [(16, 16), (32, 64), (64, 224), (8, 32)])
def test_m128_compact_layout(ctx, N, K):
    # compact: row groups 128 B apart, K chunks (M/8)*128 B apart
    dump, D = _run(ctx, 128, N, K, 128, 16 * 128, 128, (N // 8) * 128)
    assert np.array_equal(dump[:, :N], D)


def test_m128_strided_layouts(ctx):
    # the conv2 geometry: 8-row groups 160 B apart (10-wide padded rows); chunk strides chosen so that
    # the test operands do not alias (in conv2 they alias on purpose: shifted views of one volume)
    dump, D = _run(ctx, 128, 32, 16, 160, 4096, 128, 512, seed=1)
    assert np.array_equal(dump[:, :32], D)
    dump, D = _run(ctx, 128, 16, 16, 160, 3200, 256, 1024, seed=2)
    assert np.array_equal(dump[:, :16], D)
    dump, D = _run(ctx, 128, 48, 32, 288, 4800, 256, 2048, seed=3)
    assert np.array_equal(dump[:, :48], D)


@pytest.mark.parametrize("N", [16, 32])
def test_m64_lane_layout(ctx, N):
    dump, D = _run(ctx, 64, N, 64, 160, 2048, 128, (N // 8) * 128, seed=4)
    assert np.array_equal(dump[_m64_lanes(0), :N], D)


def test_m64_interleaved_second_tile(ctx):
    dump, D = _run(ctx, 64, 32, 32, 160, 2048, 128, 512, lane_off=16, seed=5)
    assert np.array_equal(dump[_m64_lanes(16), :32], D)
