// encoder.cu — a3: PatchEncoder.predict (reference Match.py:130-135; graph AE4VoxelPatch.py:189-197
// with the activations of the shipped EncoderModel4VoxelPatch.h5 = tanh on every layer).
//
//   conv3d 3^3 1->8 tanh, maxpool 2 | conv3d 8->16 tanh, maxpool 2 | conv3d 16->32 tanh |
//   flatten (x,y,z,c) | dense 2048->200 tanh | dense 200->20 tanh
//
// Input is the 512-byte bit-packed occupancy patch produced by patches.cu (values are exactly
// {0,1}); max-pool commutes with the monotonic tanh, so tanh is applied after pooling.
//
//   conv12_tc_kernel   persistent CTAs (2 per SM).  Per patch: conv1 on CUDA cores as an exact sum of
//                      selected weights (64-bit neighbourhood windows) + max-pool + tanh, written as
//                      split fp16 (hi+lo) into a zero-haloed 10^3 x 8ch volume in shared memory;
//                      conv2 as an IMPLICIT GEMM on tcgen05: for every x-slice (M = 64 positions) and
//                      tap pair (K = 16) the A operand is a shifted view of that volume described by
//                      a no-swizzle K-major smem descriptor (SBO = 160 B = one padded y-row, LBO =
//                      distance between the two taps), B = [W_hi | W_lo] (N = 32) for A_hi and W_hi
//                      (N = 16) for A_lo, fp32 accumulation in TMEM (double-buffered, 2 x 128 columns);
//                      epilogue: tcgen05.ld, hi/lo halves added, bias, 2x2x2 max-pool by warp
//                      shuffles (two interleaved M=64 tiles per 32-lane quarter), tanh -> act2.
//                      One warp issues MMAs; eight warps produce operands / drain accumulators.
//   conv3_kernel       fp32 CUDA cores (weights in smem) -> act3 [P,2048]
//   dense_kernel       64-patch tiles: dense1 (register-tiled SGEMM) + tanh + dense2 + tanh
#include "common.cuh"
#include "umma.cuh"

namespace {

// ---- conv1 (CUDA cores) + conv2 (tcgen05) --------------------------------------------------
constexpr int TC_WORKERS = 256;                       // 8 warps: produce operands (conv1), drain accumulators
constexpr int TC_THREADS = TC_WORKERS + 32;           // + 1 warp whose elected lane issues the MMAs (it
                                                      // absorbs the tensor-queue back-pressure)
constexpr int A_VOL_BYTES = 10 * 10 * 10 * 16;        // padded 8^3 volume, 8 ch fp16 per position
constexpr int W2_BYTES = 28 * 512;                    // 28 tap chunks x 32 rows x 16 B
// dynamic smem layout (bytes)
constexpr int SM_A = 0;                               // [buf 2][hi,lo][A_VOL_BYTES]
constexpr int SM_W2 = SM_A + 4 * A_VOL_BYTES;         // 64000
constexpr int SM_K1 = SM_W2 + W2_BYTES;               // floats: k1[216] b1[8] b2[16]
constexpr int SM_BG = SM_K1 + (216 + 8 + 16) * 4;     // tanh(b1) as fp16 hi (16 B) + lo (16 B)
constexpr int SM_PK = SM_BG + 32;                     // packed patch [2][128] words
constexpr int SM_BAR = SM_PK + 2 * 512;               // 6 mbarriers + tmem base
constexpr int TC_SMEM = SM_BAR + 64;

struct Conv12Args {
    const unsigned *packed;  // [P,128]
    const float *k1, *b1;    // (27,8), (8)
    const float *k2, *b2;    // (27,8,16), (16)
    float *act2;             // [P,64,16] fp32
    int P;
    long long *timeline;     // debug: [gridDim.x][64][8] clock64 stamps, or null
};

__device__ __forceinline__ constexpr int tap_off(int t)  // byte offset of tap t inside the padded volume
{
    return (((t / 9) * 10 + (t / 3) % 3) * 10 + t % 3) * 16;
}

// conv1 + maxpool + tanh for one patch -> A_hi / A_lo interior.  Thread = one (px,py) column and two
// pz cells: the 16 occupancy rows around the column are read once; a cell whose 4x4x4 window is
// empty gets the precomputed tanh(b1); otherwise the set bits are visited in ascending order and
// each adds its weight to the sub-positions it touches (same summation order as tap-ascending).
__device__ __forceinline__ void conv1_to_smem(const unsigned *pk, const float *k1s, const float *b1s,
                                              const uint4 *bg, unsigned char *a_hi, unsigned char *a_lo, int tid)
{
    const unsigned short *rows = reinterpret_cast<const unsigned short *>(pk);  // row (x,y): 16 z-bits
    const int col = tid >> 2, zq = tid & 3;
    const int px = col >> 3, py = col & 7;
    unsigned r[16];
    unsigned any = 0;
#pragma unroll
    for (int ix = 0; ix < 4; ++ix)
#pragma unroll
        for (int iy = 0; iy < 4; ++iy) {
            int x = 2 * px - 1 + ix, y = 2 * py - 1 + iy;
            unsigned v = 0;
            if (x >= 0 && x < 16 && y >= 0 && y < 16) v = rows[x * 16 + y];
            r[ix * 4 + iy] = v << 1;  // bit z+1 <-> voxel z
            any |= v;
        }
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int pz = 2 * zq + half;
        const int pi = ((px + 1) * 10 + (py + 1)) * 10 + (pz + 1);
        unsigned long long win = 0ull;
        if (any) {
#pragma unroll
            for (int i = 0; i < 16; ++i) win |= (unsigned long long)((r[i] >> (2 * pz)) & 0xFu) << (i * 4);
        }
        if (win == 0ull) {
            *reinterpret_cast<uint4 *>(a_hi + pi * 16) = bg[0];
            *reinterpret_cast<uint4 *>(a_lo + pi * 16) = bg[1];
            continue;
        }
        float acc[8][8];
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[s][c] = b1s[c];
        while (win) {
            const int bit = __ffsll((long long)win) - 1;
            win &= win - 1;
            const int ix = bit >> 4, iy = (bit >> 2) & 3, iz = bit & 3;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const int dx = ix - (s >> 2), dy = iy - ((s >> 1) & 1), dz = iz - (s & 1);
                if ((unsigned)dx < 3u && (unsigned)dy < 3u && (unsigned)dz < 3u) {
                    const float4 *w = reinterpret_cast<const float4 *>(k1s + ((dx * 3 + dy) * 3 + dz) * 8);
                    float4 w0 = w[0], w1 = w[1];
                    acc[s][0] += w0.x; acc[s][1] += w0.y; acc[s][2] += w0.z; acc[s][3] += w0.w;
                    acc[s][4] += w1.x; acc[s][5] += w1.y; acc[s][6] += w1.z; acc[s][7] += w1.w;
                }
            }
        }
        __half2 hi[4], lo[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float m0 = acc[0][2 * c], m1 = acc[0][2 * c + 1];
#pragma unroll
            for (int s = 1; s < 8; ++s) { m0 = fmaxf(m0, acc[s][2 * c]); m1 = fmaxf(m1, acc[s][2 * c + 1]); }
            __half h0, l0, h1, l1;
            umma::split_f16(tanhf(m0), h0, l0);
            umma::split_f16(tanhf(m1), h1, l1);
            hi[c] = __halves2half2(h0, h1);
            lo[c] = __halves2half2(l0, l1);
        }
        *reinterpret_cast<uint4 *>(a_hi + pi * 16) = *reinterpret_cast<uint4 *>(hi);
        *reinterpret_cast<uint4 *>(a_lo + pi * 16) = *reinterpret_cast<uint4 *>(lo);
    }
}

__global__ void __launch_bounds__(TC_THREADS, 2) conv12_tc_kernel(const Conv12Args a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float *k1s = reinterpret_cast<float *>(sm + SM_K1);
    float *b1s = k1s + 216, *b2s = b1s + 8;
    const uint4 *bg = reinterpret_cast<const uint4 *>(sm + SM_BG);
    unsigned *pk = reinterpret_cast<unsigned *>(sm + SM_PK);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + SM_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + SM_BAR + 48);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler

    // ---- one-time setup: zero the operand volumes (halo stays zero), stage weights ----
    for (int i = tid; i < (SM_K1) / 16; i += TC_THREADS) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 216; i += TC_THREADS) k1s[i] = a.k1[i];
    if (tid < 8) {
        b1s[tid] = a.b1[tid];
        __half h, l;
        umma::split_f16(tanhf(a.b1[tid]), h, l);
        reinterpret_cast<__half *>(sm + SM_BG)[tid] = h;
        reinterpret_cast<__half *>(sm + SM_BG + 16)[tid] = l;
    }
    if (tid < 16) b2s[tid] = a.b2[tid];
    __syncthreads();
    // B operand: row n (0..15 = W_hi, 16..31 = W_lo of out-channel n%16), k = tap*8 + ci; chunk = tap
    for (int e = tid; e < 27 * 8 * 16; e += TC_THREADS) {
        int t = e / 128, ci = (e / 16) % 8, co = e % 16;
        __half h, l;
        umma::split_f16(a.k2[e], h, l);
        unsigned char *w = sm + SM_W2 + t * 512;
        *reinterpret_cast<__half *>(w + (co / 8) * 128 + (co % 8) * 16 + ci * 2) = h;
        *reinterpret_cast<__half *>(w + ((16 + co) / 8) * 128 + (co % 8) * 16 + ci * 2) = l;
    }
    if (warp == 8) umma::tmem_alloc(tmem_slot, 256);
    uint64_t *full = mbar, *tfull = mbar + 2, *tempty = mbar + 4;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&full[b], TC_WORKERS);    // operands of buffer b written (every worker arrives)
            umma::mbar_init(&tfull[b], 1);            // MMAs into TMEM[b] complete (tcgen05.commit)
            umma::mbar_init(&tempty[b], TC_WORKERS);  // TMEM[b] drained by the epilogue
        }
        umma::fence_mbar_init();
    }
    umma::fence_proxy_async();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t sA = umma::smem_u32(sm + SM_A), sW = umma::smem_u32(sm + SM_W2);

    const int n_my = (a.P - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // patches of this CTA
    auto patch_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
    auto stamp = [&](int i, int slot) {
        if (a.timeline && i < 64) a.timeline[((size_t)blockIdx.x * 64 + i) * 8 + slot] = clock64();
    };

    if (warp == 8) {
        // ===== MMA issuer: waits for operands + a free accumulator, queues 224 MMAs, commits =====
        const uint32_t idesc32 = umma::idesc_f16_f32(64, 32), idesc16 = umma::idesc_f16_f32(64, 16);
        for (int j = 0; j < n_my; ++j) {
            const int b = j & 1, k = j >> 1;
            umma::mbar_wait(&full[b], (uint32_t)(k & 1));
            if (k >= 1) umma::mbar_wait(&tempty[b], (uint32_t)((k - 1) & 1));
            umma::fence_after_thread_sync();
            if (lane == 0) stamp(j, 5);
            if (umma::elect_one()) {
#pragma unroll 1
                for (int xs = 0; xs < 8; ++xs) {  // x-slice: M = 64 positions (y,z)
                    const uint32_t d = tbase + ((uint32_t)((xs & 1) * 16) << 16) + b * 128 + (xs >> 1) * 32;
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        const uint32_t abase = sA + (2 * b + part) * A_VOL_BYTES + xs * 1600;
#pragma unroll
                        for (int s = 0; s < 14; ++s) {
                            const int t0 = 2 * s;
                            const int lbo = (s == 13) ? 16 : tap_off(t0 + 1) - tap_off(t0);
                            uint64_t da = umma::smem_desc(abase + tap_off(t0), lbo, 160);
                            uint64_t db = umma::smem_desc(sW + s * 1024, 512, 128);
                            umma::mma_f16(d, da, db, part ? idesc16 : idesc32, (part | s) ? 1u : 0u);
                        }
                    }
                }
                umma::commit(&tfull[b]);
            }
            __syncwarp();
            if (lane == 0) stamp(j, 6);
        }
    } else {
        // ===== workers: conv1 of patch i+1 while MMA(i) runs, then drain patch i =====
        auto produce = [&](int i) {
            const int b = i & 1;
            if (tid < 128) pk[b * 128 + tid] = a.packed[(size_t)patch_of(i) * 128 + tid];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            conv1_to_smem(pk + b * 128, k1s, b1s, bg, sm + SM_A + (2 * b) * A_VOL_BYTES,
                          sm + SM_A + (2 * b + 1) * A_VOL_BYTES, tid);
            umma::fence_proxy_async();
            umma::mbar_arrive(&full[b]);
        };
        if (n_my > 0) produce(0);
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            if (tid == 0) stamp(i, 0);
            if (i + 1 < n_my) produce(i + 1);  // MMA(i-1) finished reading A[b^1]: waited on tfull last iteration
            if (tid == 0) stamp(i, 1);
            umma::mbar_wait(&tfull[b], (uint32_t)((i >> 1) & 1));
            umma::fence_after_thread_sync();
            if (tid == 0) stamp(i, 3);
            const int q = warp & 3, h = warp >> 2;
            float *out = a.act2 + (size_t)patch_of(i) * 1024;
#pragma unroll 1
            for (int pp = 0; pp < 2; ++pp) {
                const int pair = 2 * h + pp;
                uint32_t v[32];
                umma::tmem_ld_x32(tbase + ((uint32_t)(32 * q) << 16) + b * 128 + pair * 32, v);
                umma::tmem_ld_wait();
                float m[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) m[c] = (__uint_as_float(v[c]) + __uint_as_float(v[16 + c])) + b2s[c];
                // 2x2x2 max-pool: partners differ in z (lane^1), y (lane^8), x-slice (lane^16)
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    m[c] = fmaxf(m[c], __shfl_xor_sync(0xffffffffu, m[c], 1));
                    m[c] = fmaxf(m[c], __shfl_xor_sync(0xffffffffu, m[c], 8));
                    m[c] = fmaxf(m[c], __shfl_xor_sync(0xffffffffu, m[c], 16));
                }
                // the 8 lanes of a pooling group share the result; lane j of the group finishes channels 2j, 2j+1
                const int j = (lane & 1) | ((lane >> 2) & 2) | ((lane >> 2) & 4);
                float o0 = 0.f, o1 = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c == j) { o0 = m[2 * c]; o1 = m[2 * c + 1]; }
                const int pos = (pair * 4 + q) * 4 + ((lane & 7) >> 1);
                *reinterpret_cast<float2 *>(out + pos * 16 + 2 * j) = make_float2(tanhf(o0), tanhf(o1));
            }
            umma::fence_before_thread_sync();
            umma::mbar_arrive(&tempty[b]);
            if (tid == 0) stamp(i, 4);
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    if (warp == 8) umma::tmem_dealloc(tbase, 256);
}

// ---- conv3 (fp32 CUDA cores) ------------------------------------------------------------------
constexpr int C3_THREADS = 256;
constexpr int A2P = 6 * 6 * 6;  // padded 4^3 volume
constexpr int C3_SMEM = (27 * 16 * 32 + A2P * 16 + 32) * 4;

struct Conv3Args {
    const float *act2;     // [P,64,16]
    const float *k3, *b3;  // (27,16,32), (32)
    float *act3;           // [P,2048]
    int P;
};

__global__ void __launch_bounds__(C3_THREADS, 2) conv3_kernel(const Conv3Args a)
{
    extern __shared__ __align__(16) float sm3[];
    float *k3s = sm3, *a2 = sm3 + 27 * 16 * 32, *b3s = a2 + A2P * 16;
    const int tid = threadIdx.x;
    for (int i = tid; i < 27 * 16 * 32; i += C3_THREADS) k3s[i] = a.k3[i];
    for (int i = tid; i < A2P * 16; i += C3_THREADS) a2[i] = 0.0f;
    if (tid < 32) b3s[tid] = a.b3[tid];
    __syncthreads();
    for (int p = blockIdx.x; p < a.P; p += gridDim.x) {
        {   // 64 positions x 16 ch = 256 float4
            int pos = tid >> 2, c4 = tid & 3;
            int x = pos >> 4, y = (pos >> 2) & 3, z = pos & 3;
            float4 v = *reinterpret_cast<const float4 *>(a.act2 + (size_t)p * 1024 + pos * 16 + c4 * 4);
            *reinterpret_cast<float4 *>(a2 + (((x + 1) * 6 + (y + 1)) * 6 + (z + 1)) * 16 + c4 * 4) = v;
        }
        __syncthreads();
        {
            const int g = tid & 7, qq = tid >> 3;  // channels 4g..4g+3, positions qq and qq+32
            float acc[2][4];
            int base[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int q = qq + 32 * j;
                int x = q >> 4, y = (q >> 2) & 3, z = q & 3;
                base[j] = (x * 6 + y) * 6 + z;
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[j][c] = b3s[g * 4 + c];
            }
            for (int t = 0; t < 27; ++t) {
                const int toff = ((t / 9) * 6 + (t / 3) % 3) * 6 + t % 3;
                const float4 *i0 = reinterpret_cast<const float4 *>(a2 + (base[0] + toff) * 16);
                const float4 *i1 = reinterpret_cast<const float4 *>(a2 + (base[1] + toff) * 16);
                const float *wt = k3s + t * 512 + g * 4;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    float4 u0 = i0[c4], u1 = i1[c4];
                    float x0[4] = {u0.x, u0.y, u0.z, u0.w}, x1[4] = {u1.x, u1.y, u1.z, u1.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float4 w = *reinterpret_cast<const float4 *>(wt + (c4 * 4 + k) * 32);
                        acc[0][0] = fmaf(x0[k], w.x, acc[0][0]); acc[0][1] = fmaf(x0[k], w.y, acc[0][1]);
                        acc[0][2] = fmaf(x0[k], w.z, acc[0][2]); acc[0][3] = fmaf(x0[k], w.w, acc[0][3]);
                        acc[1][0] = fmaf(x1[k], w.x, acc[1][0]); acc[1][1] = fmaf(x1[k], w.y, acc[1][1]);
                        acc[1][2] = fmaf(x1[k], w.z, acc[1][2]); acc[1][3] = fmaf(x1[k], w.w, acc[1][3]);
                    }
                }
            }
            float *o = a.act3 + (size_t)p * 2048;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int q = qq + 32 * j;  // flatten index ((x*4+y)*4+z)*32 + c
                *reinterpret_cast<float4 *>(o + q * 32 + g * 4) =
                    make_float4(tanhf(acc[j][0]), tanhf(acc[j][1]), tanhf(acc[j][2]), tanhf(acc[j][3]));
            }
        }
        __syncthreads();
    }
}

// ---- dense1 + tanh + dense2 + tanh ---------------------------------------------------------
constexpr int DT_ROWS = 64, DT_K = 16, DT_COLS = 208;  // 200 padded to 16*13
constexpr int DT_SMEM = (DT_K * (DT_ROWS + 4) + DT_K * DT_COLS + DT_ROWS * 201) * 4;
struct DenseArgs {
    const float *act3;  // [P,2048]
    const float *d1, *bd1, *d2, *bd2;
    float *feat;
    int P, feat_stride, feat_col0;
    // frame mode: packed order is [F,3,K]; row p -> feat[(f*K+k)*60 + s*20]
    int frame_mode, K;
};

__global__ void __launch_bounds__(256) dense_kernel(const DenseArgs a)
{
    extern __shared__ __align__(16) float dsm[];
    float (*As)[DT_ROWS + 4] = reinterpret_cast<float (*)[DT_ROWS + 4]>(dsm);
    float (*Bs)[DT_COLS] = reinterpret_cast<float (*)[DT_COLS]>(dsm + DT_K * (DT_ROWS + 4));
    float (*Hs)[201] = reinterpret_cast<float (*)[201]>(dsm + DT_K * (DT_ROWS + 4) + DT_K * DT_COLS);
    const int tid = threadIdx.x;
    const int tr = tid >> 4, tc = tid & 15;  // 16 x 16 thread grid: 4 rows x 13 cols each
    const int row0 = blockIdx.x * DT_ROWS;
    float acc[4][13];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 13; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < 2048; k0 += DT_K) {
        // A tile: 64 rows x 16 k  (one float4 per thread)
        {
            int r = tid >> 2, kk = (tid & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < a.P) v = *reinterpret_cast<const float4 *>(a.act3 + (size_t)(row0 + r) * 2048 + k0 + kk);
            As[kk + 0][r] = v.x; As[kk + 1][r] = v.y; As[kk + 2][r] = v.z; As[kk + 3][r] = v.w;
        }
        // B tile: 16 k x 200 cols
        for (int i = tid; i < DT_K * 50; i += 256) {
            int kk = i / 50, c4 = i % 50;
            float4 v = *reinterpret_cast<const float4 *>(a.d1 + (size_t)(k0 + kk) * 200 + c4 * 4);
            *reinterpret_cast<float4 *>(&Bs[kk][c4 * 4]) = v;
        }
        if (tid < DT_K * 2) {
            int kk = tid >> 1, c4 = 50 + (tid & 1);
            *reinterpret_cast<float4 *>(&Bs[kk][c4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < DT_K; ++kk) {
            float av[4], bv[13];
            float4 a4 = *reinterpret_cast<const float4 *>(&As[kk][tr * 4]);
            av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
#pragma unroll
            for (int j = 0; j < 13; ++j) bv[j] = Bs[kk][tc + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 13; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            int c = tc + 16 * j;
            if (c < 200) Hs[tr * 4 + i][c] = tanhf(acc[i][j] + a.bd1[c]);
        }
    __syncthreads();
    // dense2: 64 rows x 20 outputs = 1280 = 5 per thread
    for (int o = tid; o < DT_ROWS * 20; o += 256) {
        int r = o / 20, c = o % 20;
        int p = row0 + r;
        if (p >= a.P) continue;
        float s = a.bd2[c];
        for (int k = 0; k < 200; ++k) s = fmaf(Hs[r][k], __ldg(a.d2 + k * 20 + c), s);
        float v = tanhf(s);
        if (a.frame_mode) {
            int k = p % a.K, sc = (p / a.K) % 3, f = p / (3 * a.K);
            a.feat[((size_t)f * a.K + k) * 60 + sc * 20 + c] = v;
        } else {
            a.feat[(size_t)p * a.feat_stride + a.feat_col0 + c] = v;
        }
    }
}

// ---- f32 patches -> packed bits (the predict() boundary) ------------------------------------
__global__ void pack_kernel(const float *__restrict__ patches, unsigned *__restrict__ packed,
                            long long nwords, int *status)
{
    // one warp per packed word: lane l reads float l of the 32 (coalesced), ballot packs
    const int lane = threadIdx.x & 31;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (long long w = warp; w < nwords; w += nwarps) {
        float v = patches[w * 32 + lane];
        bad |= !(v == 0.0f || v == 1.0f);
        unsigned m = __ballot_sync(0xffffffffu, v != 0.0f);
        if (lane == 0) packed[w] = m;
    }
    if (bad && status) atomicExch(status, CAELO_ERR_NONBINARY_PATCH);
}

int run_encoder(caelo_ctx *ctx, const unsigned *packed, int P, float *feat, int feat_stride,
                int feat_col0, int frame_mode, int K, cudaStream_t st)
{
    if (!ctx->have_encoder) return CAELO_ERR_NO_WEIGHTS;
    if (P <= 0) return CAELO_OK;
    int rc = caelo_reserve(ctx, ctx->enc_ws, (size_t)P * (2048 + 1024) * 4);
    if (rc) return rc;
    float *act3 = reinterpret_cast<float *>(ctx->enc_ws.ptr);
    float *act2 = act3 + (size_t)P * 2048;
    Conv12Args c;
    c.packed = packed; c.k1 = ctx->enc.k1; c.b1 = ctx->enc.b1; c.k2 = ctx->enc.k2; c.b2 = ctx->enc.b2;
    c.act2 = act2; c.P = P; c.timeline = ctx->dbg_timeline;
    int grid = 2 * ctx->num_sms < P ? 2 * ctx->num_sms : P;
    { ProfScope ps_(ctx, "conv12_tc_kernel", st); conv12_tc_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(c); }
    CAELO_LAUNCH_CHECK(ctx);
    Conv3Args c3;
    c3.act2 = act2; c3.k3 = ctx->enc.k3; c3.b3 = ctx->enc.b3; c3.act3 = act3; c3.P = P;
    { ProfScope ps_(ctx, "conv3_kernel", st); conv3_kernel<<<grid, C3_THREADS, C3_SMEM, st>>>(c3); }
    CAELO_LAUNCH_CHECK(ctx);
    DenseArgs d;
    d.act3 = act3; d.d1 = ctx->enc.d1; d.bd1 = ctx->enc.bd1; d.d2 = ctx->enc.d2; d.bd2 = ctx->enc.bd2;
    d.feat = feat; d.P = P; d.feat_stride = feat_stride; d.feat_col0 = feat_col0;
    d.frame_mode = frame_mode; d.K = K;
    { ProfScope ps_(ctx, "dense_kernel", st); dense_kernel<<<(P + DT_ROWS - 1) / DT_ROWS, 256, DT_SMEM, st>>>(d); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}

}  // namespace

int caelo_encoder_init(caelo_ctx *ctx)
{
    CAELO_CUDA(ctx, cudaFuncSetAttribute(conv12_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(conv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM));
    return CAELO_OK;
}

extern "C" int caelo_encode_packed(caelo_ctx *ctx, const uint32_t *packed, int P, float *feat,
                                   int feat_stride, int feat_col0, void *stream)
{
    if (!ctx || !packed || !feat || P < 0 || feat_stride < 20 || feat_col0 < 0) return CAELO_ERR_ARG;
    return run_encoder(ctx, packed, P, feat, feat_stride, feat_col0, 0, 1, (cudaStream_t)stream);
}

extern "C" int caelo_encode_frames(caelo_ctx *ctx, const uint32_t *packed, int F, int K, float *feat,
                                   void *stream)
{
    if (!ctx || !packed || !feat || F <= 0 || K <= 0) return CAELO_ERR_ARG;
    return run_encoder(ctx, packed, F * 3 * K, feat, 60, 0, 1, K, (cudaStream_t)stream);
}

extern "C" int caelo_encode_patches(caelo_ctx *ctx, const float *patches, int P, float *feat,
                                    int32_t *status, void *stream)
{
    if (!ctx || !patches || !feat || P < 0) return CAELO_ERR_ARG;
    if (P == 0) return CAELO_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = caelo_reserve(ctx, ctx->misc, (size_t)P * 512);
    if (rc) return rc;
    unsigned *packed = reinterpret_cast<unsigned *>(ctx->misc.ptr);
    if (status) CAELO_CUDA(ctx, cudaMemsetAsync(status, 0, 4, st));
    long long nwords = (long long)P * 128;
    long long blocks = (nwords * 32 + 255) / 256;
    if (blocks > (long long)ctx->num_sms * 32) blocks = (long long)ctx->num_sms * 32;
    { ProfScope ps_(ctx, "pack_kernel", st); pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(patches, packed, nwords, status); }
    CAELO_LAUNCH_CHECK(ctx);
    return run_encoder(ctx, packed, P, feat, 20, 0, 0, 1, st);
}
