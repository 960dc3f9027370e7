// encoder.cu — a3: PatchEncoder.predict (reference Match.py:130-135; graph AE4VoxelPatch.py:189-197
// with the activations of the shipped EncoderModel4VoxelPatch.h5 = tanh on every layer).
//
//   conv3d 3^3 1->8 tanh, maxpool 2 | conv3d 8->16 tanh, maxpool 2 | conv3d 16->32 tanh |
//   flatten (x,y,z,c) | dense 2048->200 tanh | dense 200->20 tanh
//
// Input is the 512-byte bit-packed occupancy patch produced by patches.cu (values are exactly
// {0,1}); max-pool commutes with the monotonic tanh, so tanh is applied after pooling.
//
// v1 (this file): fp32 CUDA-core path.
//   conv_stack_kernel  persistent CTAs, conv2/conv3 weights resident in shared memory; per patch
//                      conv1 as an exact sum of selected weights (64-bit neighbourhood windows),
//                      conv2, conv3 -> act3 [P,2048] in HBM/L2
//   dense_kernel       64-patch tiles: dense1 (register-tiled SGEMM) + tanh + dense2 + tanh
#include "common.cuh"

namespace {

constexpr int CS_THREADS = 256;

struct ConvArgs {
    const unsigned *packed;  // [P,128]
    const float *k1, *b1;    // (27,8), (8)
    const float *k2, *b2;    // (27,8,16), (16)
    const float *k3, *b3;    // (27,16,32), (32)
    float *act3;             // [P,2048]
    int P;
};

// shared-memory layout (floats)
constexpr int A1P = 10 * 10 * 10;       // padded 8^3 volume
constexpr int A2P = 6 * 6 * 6;          // padded 4^3 volume
constexpr int OFF_K2 = 0;                          // 27*8*16   = 3456
constexpr int OFF_K3 = OFF_K2 + 27 * 8 * 16;       // 27*16*32  = 13824
constexpr int OFF_A1 = OFF_K3 + 27 * 16 * 32;      // [2][A1P][4]
constexpr int OFF_C2 = OFF_A1 + 2 * A1P * 4;       // [512][16] conv2 pre-activation
constexpr int OFF_A2 = OFF_C2 + 512 * 16;          // [A2P][16]
constexpr int OFF_K1 = OFF_A2 + A2P * 16;          // 27*8 + 8 + 16 + 32 (k1,b1,b2,b3)
constexpr int OFF_PK = OFF_K1 + 27 * 8 + 8 + 16 + 32;  // 128 words packed patch
constexpr int CS_SMEM_FLOATS = OFF_PK + 128;

__global__ void __launch_bounds__(CS_THREADS, 1) conv_stack_kernel(const ConvArgs a)
{
    extern __shared__ __align__(16) float sm[];
    float *k2s = sm + OFF_K2, *k3s = sm + OFF_K3, *a1 = sm + OFF_A1, *c2 = sm + OFF_C2,
          *a2 = sm + OFF_A2, *k1s = sm + OFF_K1;
    float *b1s = k1s + 27 * 8, *b2s = b1s + 8, *b3s = b2s + 16;
    unsigned *pk = reinterpret_cast<unsigned *>(sm + OFF_PK);
    const int tid = threadIdx.x;

    for (int i = tid; i < 27 * 8 * 16; i += CS_THREADS) k2s[i] = a.k2[i];
    for (int i = tid; i < 27 * 16 * 32; i += CS_THREADS) k3s[i] = a.k3[i];
    for (int i = tid; i < 27 * 8; i += CS_THREADS) k1s[i] = a.k1[i];
    if (tid < 8) b1s[tid] = a.b1[tid];
    if (tid < 16) b2s[tid] = a.b2[tid];
    if (tid < 32) b3s[tid] = a.b3[tid];
    for (int i = tid; i < 2 * A1P * 4; i += CS_THREADS) a1[i] = 0.0f;  // zero halo, kept for all patches
    for (int i = tid; i < A2P * 16; i += CS_THREADS) a2[i] = 0.0f;
    __syncthreads();

    for (int p = blockIdx.x; p < a.P; p += gridDim.x) {
        if (tid < 128) pk[tid] = a.packed[(size_t)p * 128 + tid];
        __syncthreads();

        // ---- conv1 (1->8) + maxpool + tanh : 512 pooled positions, 2 per thread ----
        const unsigned short *rows = reinterpret_cast<const unsigned short *>(pk);  // row (x,y): 16 z-bits
        for (int q = tid; q < 512; q += CS_THREADS) {
            const int px = q >> 6, py = (q >> 3) & 7, pz = q & 7;
            // 4x4x4 occupancy window covering the 2x2x2 pooling cell plus a halo of one
            unsigned long long win = 0ull;
#pragma unroll
            for (int ix = 0; ix < 4; ++ix)
#pragma unroll
                for (int iy = 0; iy < 4; ++iy) {
                    int x = 2 * px - 1 + ix, y = 2 * py - 1 + iy;
                    unsigned r = 0;
                    if (x >= 0 && x < 16 && y >= 0 && y < 16) r = rows[x * 16 + y];
                    unsigned four = ((r << 1) >> (2 * pz)) & 0xFu;  // z = 2pz-1 .. 2pz+2
                    win |= (unsigned long long)four << ((ix * 4 + iy) * 4);
                }
            float best[8];
            if (win == 0ull) {
#pragma unroll
                for (int c = 0; c < 8; ++c) best[c] = b1s[c];
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) best[c] = -3.0e38f;
                for (int sx = 0; sx < 2; ++sx)
                    for (int sy = 0; sy < 2; ++sy)
                        for (int sz = 0; sz < 2; ++sz) {
                            float acc[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) acc[c] = b1s[c];
#pragma unroll
                            for (int t = 0; t < 27; ++t) {
                                int dx = t / 9, dy = (t / 3) % 3, dz = t % 3;
                                int bit = (((sx + dx) * 4 + (sy + dy)) * 4) + (sz + dz);
                                if ((win >> bit) & 1ull) {
#pragma unroll
                                    for (int c = 0; c < 8; ++c) acc[c] += k1s[t * 8 + c];
                                }
                            }
#pragma unroll
                            for (int c = 0; c < 8; ++c) best[c] = fmaxf(best[c], acc[c]);
                        }
            }
            const int pi = ((px + 1) * 10 + (py + 1)) * 10 + (pz + 1);
            *reinterpret_cast<float4 *>(a1 + pi * 4) =
                make_float4(tanhf(best[0]), tanhf(best[1]), tanhf(best[2]), tanhf(best[3]));
            *reinterpret_cast<float4 *>(a1 + (A1P + pi) * 4) =
                make_float4(tanhf(best[4]), tanhf(best[5]), tanhf(best[6]), tanhf(best[7]));
        }
        __syncthreads();

        // ---- conv2 (8->16): 512 positions, 2 per thread, 16 channels each ----
        {
            float acc[2][16];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[j][c] = b2s[c];
            int base[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int q = tid + j * CS_THREADS;
                int x = q >> 6, y = (q >> 3) & 7, z = q & 7;
                base[j] = (x * 10 + y) * 10 + z;  // padded index of tap (0,0,0)
            }
            for (int t = 0; t < 27; ++t) {
                const int toff = ((t / 9) * 10 + (t / 3) % 3) * 10 + t % 3;
                float in[2][8];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float4 lo = *reinterpret_cast<const float4 *>(a1 + (base[j] + toff) * 4);
                    float4 hi = *reinterpret_cast<const float4 *>(a1 + (A1P + base[j] + toff) * 4);
                    in[j][0] = lo.x; in[j][1] = lo.y; in[j][2] = lo.z; in[j][3] = lo.w;
                    in[j][4] = hi.x; in[j][5] = hi.y; in[j][6] = hi.z; in[j][7] = hi.w;
                }
                const float4 *w4 = reinterpret_cast<const float4 *>(k2s + t * 128);
#pragma unroll
                for (int ci = 0; ci < 8; ++ci) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float4 w = w4[ci * 4 + g];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            acc[j][g * 4 + 0] = fmaf(in[j][ci], w.x, acc[j][g * 4 + 0]);
                            acc[j][g * 4 + 1] = fmaf(in[j][ci], w.y, acc[j][g * 4 + 1]);
                            acc[j][g * 4 + 2] = fmaf(in[j][ci], w.z, acc[j][g * 4 + 2]);
                            acc[j][g * 4 + 3] = fmaf(in[j][ci], w.w, acc[j][g * 4 + 3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int q = tid + j * CS_THREADS;
                // channel-group-major so that the pooling reads are conflict-free: c2[g][q][4]
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<float4 *>(c2 + (g * 512 + q) * 4) =
                        make_float4(acc[j][g * 4], acc[j][g * 4 + 1], acc[j][g * 4 + 2], acc[j][g * 4 + 3]);
            }
        }
        __syncthreads();

        // ---- maxpool + tanh -> a2 padded [6,6,6][16]: 64 positions x 4 channel groups ----
        {
            const int g = tid & 3, q = tid >> 2;  // q: pooled position 0..63
            const int px = q >> 4, py = (q >> 2) & 3, pz = q & 3;
            float4 m = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                int x = 2 * px + (s >> 2), y = 2 * py + ((s >> 1) & 1), z = 2 * pz + (s & 1);
                float4 v = *reinterpret_cast<const float4 *>(c2 + (g * 512 + (x * 8 + y) * 8 + z) * 4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
            const int pi = ((px + 1) * 6 + (py + 1)) * 6 + (pz + 1);
            *reinterpret_cast<float4 *>(a2 + pi * 16 + g * 4) =
                make_float4(tanhf(m.x), tanhf(m.y), tanhf(m.z), tanhf(m.w));
        }
        __syncthreads();

        // ---- conv3 (16->32) + tanh: thread = 2 positions x 4 channels ----
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
    int rc = caelo_reserve(ctx, ctx->enc_ws, (size_t)P * 2048 * 4);
    if (rc) return rc;
    float *act3 = reinterpret_cast<float *>(ctx->enc_ws.ptr);
    ConvArgs c;
    c.packed = packed; c.k1 = ctx->enc.k1; c.b1 = ctx->enc.b1; c.k2 = ctx->enc.k2; c.b2 = ctx->enc.b2;
    c.k3 = ctx->enc.k3; c.b3 = ctx->enc.b3; c.act3 = act3; c.P = P;
    int grid = ctx->num_sms < P ? ctx->num_sms : P;
    { ProfScope ps_(ctx, "conv_stack_kernel", st); conv_stack_kernel<<<grid, CS_THREADS, CS_SMEM_FLOATS * 4, st>>>(c); }
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
    CAELO_CUDA(ctx, cudaFuncSetAttribute(conv_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         CS_SMEM_FLOATS * 4));
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
