// match.cu — a4: descriptor nearest-neighbour match (reference Match.py:257-258:
// cdist(Codes0, Codes1,'euclidean') in float64, then argmin(axis=0), ties -> lowest row).
//
// Index-exact in two steps:
//   nn_tile_kernel   float32 direct-difference distances on 64x64 tiles held in shared memory;
//                    per column the best and second-best approximate d^2 (+ best row).  The
//                    float32 sum of D non-negative terms is within (D+4)*2^-24 relative of the
//                    true value, so a column whose runner-up is outside that margin is decided.
//   nn_exact_kernel  one warp per column; undecided columns are re-scanned with the reference's
//                    own arithmetic (contract M1: float64 sequential sum, sqrt, lowest row wins).
#include "common.cuh"

namespace {

constexpr int TILE = 64;
constexpr int MT_THREADS = 256;

struct NNArgs {
    const float *c0, *c1;  // [P,N,D], [P,M,D]
    int N, M, D, Dp;       // Dp = D rounded up to 4
    float *best_d;         // [P,M]
    float *second_d;       // [P,M]
    int *best_i;           // [P,M]
    long long *out;        // [P,M]
};

__global__ void __launch_bounds__(MT_THREADS) nn_tile_kernel(const NNArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float *Bs = smem;                        // [Dp][TILE] columns (frame-1 descriptors), k-major
    float *As = smem + (size_t)a.Dp * TILE;  // [Dp][TILE] rows (frame-0 descriptors)
    __shared__ float r_best[16][TILE], r_second[16][TILE];
    __shared__ int r_idx[16][TILE];
    const int pair = blockIdx.y;
    const int j0 = blockIdx.x * TILE;
    const int tid = threadIdx.x, tj = tid & 15, ti = tid >> 4;
    const float *c0 = a.c0 + (size_t)pair * a.N * a.D;
    const float *c1 = a.c1 + (size_t)pair * a.M * a.D;
    const float INF = __int_as_float(0x7f800000);

    // transposing loads: consecutive threads take consecutive tile rows (conflict-free stores;
    // the strided global reads are absorbed by L1)
    for (int e = tid; e < TILE * a.Dp; e += MT_THREADS) {
        int c = e % TILE, k = e / TILE;
        float v = 0.0f;
        if (j0 + c < a.M && k < a.D) v = c1[(size_t)(j0 + c) * a.D + k];
        Bs[k * TILE + c] = v;
    }
    float best[4], second[4];
    int besti[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { best[c] = INF; second[c] = INF; besti[c] = 0x7fffffff; }

    for (int i0 = 0; i0 < a.N; i0 += TILE) {
        __syncthreads();
        for (int e = tid; e < TILE * a.Dp; e += MT_THREADS) {
            int r = e % TILE, k = e / TILE;
            float v = 0.0f;
            if (i0 + r < a.N && k < a.D) v = c0[(size_t)(i0 + r) * a.D + k];
            As[k * TILE + r] = v;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
#pragma unroll 4
        for (int k = 0; k < a.Dp; ++k) {
            float4 av = *reinterpret_cast<const float4 *>(As + k * TILE + ti * 4);
            float4 bv = *reinterpret_cast<const float4 *>(Bs + k * TILE + tj * 4);
            float ar[4] = {av.x, av.y, av.z, av.w}, bc[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float d = ar[r] - bc[c];
                    acc[r][c] = fmaf(d, d, acc[r][c]);
                }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i = i0 + ti * 4 + r;
            if (i < a.N) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float d = acc[r][c];
                    if (d < best[c]) { second[c] = best[c]; best[c] = d; besti[c] = i; }
                    else if (d < second[c]) second[c] = d;
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        r_best[ti][tj * 4 + c] = best[c];
        r_second[ti][tj * 4 + c] = second[c];
        r_idx[ti][tj * 4 + c] = besti[c];
    }
    __syncthreads();
    if (tid < TILE && j0 + tid < a.M) {
        float b = INF, s = INF;
        int bi = 0x7fffffff;
        for (int t = 0; t < 16; ++t) {
            float d = r_best[t][tid], s2 = r_second[t][tid];
            int i = r_idx[t][tid];
            if (d < b || (d == b && i < bi)) { s = fminf(s, b); b = d; bi = i; }
            else s = fminf(s, d);
            s = fminf(s, s2);
        }
        size_t o = (size_t)pair * a.M + j0 + tid;
        a.best_d[o] = b; a.second_d[o] = s; a.best_i[o] = bi;
    }
}

__global__ void __launch_bounds__(256) nn_exact_kernel(const NNArgs a, int P)
{
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (long long)P * a.M) return;
    const int pair = (int)(w / a.M), j = (int)(w % a.M);
    const float b = a.best_d[w], s = a.second_d[w];
    // relative error of each float32 d^2 <= (D+4)*2^-24; undecided if the intervals can overlap
    const float eps = (float)(a.D + 4) * 5.9604645e-8f;
    const bool decided = s > b * (1.0f + 4.0f * eps) + 1e-30f;
    if (decided) {
        if (lane == 0) a.out[w] = a.best_i[w];
        return;
    }
    const float *c0 = a.c0 + (size_t)pair * a.N * a.D;
    const float *q = a.c1 + ((size_t)pair * a.M + j) * a.D;
    double bd = __longlong_as_double(0x7ff0000000000000ll);
    int bi = 0x7fffffff;
    for (int i = lane; i < a.N; i += 32) {
        const float *p = c0 + (size_t)i * a.D;
        double acc = 0.0;
        for (int k = 0; k < a.D; ++k) {
            double d = __dsub_rn((double)p[k], (double)q[k]);
            acc = __dadd_rn(acc, __dmul_rn(d, d));
        }
        double dist = __dsqrt_rn(acc);
        if (dist < bd) { bd = dist; bi = i; }  // ascending i within a lane: first minimum kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double od = __shfl_xor_sync(0xffffffffu, bd, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (lane == 0) a.out[w] = bi;
}

}  // namespace

int caelo_match_init(caelo_ctx *ctx)
{
    CAELO_CUDA(ctx, cudaFuncSetAttribute(nn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         2 * 256 * TILE * 4));
    return CAELO_OK;
}

extern "C" int caelo_nn_match(caelo_ctx *ctx, const float *codes0, const float *codes1, int P, int N,
                              int M, int D, int64_t *pair_idx, void *stream)
{
    if (!ctx || !codes0 || !codes1 || !pair_idx || P <= 0 || N <= 0 || M <= 0 || D <= 0 || D > 256)
        return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    size_t cols = (size_t)P * M;
    int rc = caelo_reserve(ctx, ctx->misc, cols * 12 + 256);
    if (rc) return rc;
    NNArgs a;
    a.c0 = codes0; a.c1 = codes1; a.N = N; a.M = M; a.D = D; a.Dp = (D + 3) & ~3;
    a.best_d = reinterpret_cast<float *>(ctx->misc.ptr);
    a.second_d = a.best_d + cols;
    a.best_i = reinterpret_cast<int *>(a.second_d + cols);
    a.out = reinterpret_cast<long long *>(pair_idx);
    dim3 grid((M + TILE - 1) / TILE, P);
    size_t smem = (size_t)2 * a.Dp * TILE * 4;
    { ProfScope ps_(ctx, "nn_tile_kernel", st); nn_tile_kernel<<<grid, MT_THREADS, smem, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    long long threads = (long long)cols * 32;
    { ProfScope ps_(ctx, "nn_exact_kernel", st); nn_exact_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(a, P); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}
