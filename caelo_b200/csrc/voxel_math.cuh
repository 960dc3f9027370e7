// voxel_math.cuh — point -> voxel indices exactly as Voxelization computes them (reference Voxel.py:89-148),
// contract V1 of oracle/caelo_oracle.c: float64 on float32 inputs with explicit rounding intrinsics.
#pragma once

struct VoxelOfPoint {
    int b[3];    // 1.28 m block (iBlockX, iBlockY, iBlockZ)
    int g0[3];   // 2 cm voxel, via the BLOCK route: int((x_ - iBlock*1.28) / 0.02) + iBlock*64  (:120-139)
    int g1[3];   // 16 cm voxel: int(x_ / 0.16)  (:143-145)
    int g2[3];   // 64 cm voxel: int(x_ / 0.64)  (:146-148)
};

// Voxel.py:15-52: nBlocksL = int(200/1.28) = 156, nBlocksH = int(30/1.28) = 23, Visible* = nBlocks/2 * 1.28
__device__ __forceinline__ double voxel_visible(int c) { return c < 2 ? 156 / 2.0 * 1.28 : 23 / 2.0 * 1.28; }

// +1: voxel indices valid; 0: filtered by FilterOutTooFarPts (:89-97); -1: indexes outside the block grid
// (the reference raises IndexError there).
__device__ __forceinline__ int voxel_of_point(float fx, float fy, float fz, VoxelOfPoint &v)
{
    const double brs = 1.28, vs0 = 0.02, vs1 = 0.02 * 8, vs2 = 0.02 * 32;
    const float p[3] = {fx, fy, fz};
    const int nb[3] = {156, 156, 23};
#pragma unroll
    for (int c = 0; c < 3; ++c)
        if (fabs((double)p[c]) > voxel_visible(c)) return 0;
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double x_ = __dadd_rn((double)p[c], voxel_visible(c));
        const int b = (int)__ddiv_rn(x_, brs);
        const int l = (int)__ddiv_rn(__dsub_rn(x_, __dmul_rn((double)b, brs)), vs0);
        ok = ok && b >= 0 && b < nb[c] && l >= 0 && l < 64;
        v.b[c] = b;
        v.g0[c] = l + b * 64;
        v.g1[c] = (int)__ddiv_rn(x_, vs1);
        v.g2[c] = (int)__ddiv_rn(x_, vs2);
    }
    return ok ? 1 : -1;
}
