// icp.cu — f4: the device side of the ICP refinement on extended key points (reference MyICP.py:28-73 `ICP`,
// :76-85 `GetPtsInliners`; called from RefinePoses.py:273-334).
//
// COMPILED WITH -fmad=false (contracts N1 and U1 of oracle/oracle.py are plain float64 +,-,*,sqrt).
//
//   nn3_kernel        exact 1-nearest neighbour of every PC1 point among PC0 — what the reference asks sklearn's
//                     kd-tree for, by brute force: float32 points widened to float64,
//                     d = sqrt(((dx*dx) + (dy*dy)) + (dz*dz)), ties -> lowest PC0 index; one thread per query,
//                     PC0 streamed through shared memory as float64 tiles; optional inlier mask (d < thr) + count
//   transform_kernel  PC1 <- R PC1 + T with float64 products and sums, one rounding to float32 (MyICP.py:51)
// The iteration loop itself (threshold decay, Euler-angle test, R*/T* accumulation in float64) is host code in
// caelo_b200/api.py, as it is in the reference; SolveRT on the inliers is caelo_kabsch (pose.cu).
#include "common.cuh"

namespace {

constexpr int NN3_THREADS = 128;
constexpr int NN3_TILE = 1024;

__global__ void __launch_bounds__(NN3_THREADS) nn3_kernel(const float *__restrict__ pc0, int N, const float *__restrict__ pc1,
                                                          int M, long long *idx, double *dist, double thr,
                                                          unsigned char *mask, int *count)
{
    __shared__ double tile[NN3_TILE * 3];
    const int j = blockIdx.x * NN3_THREADS + threadIdx.x;
    double qx = 0.0, qy = 0.0, qz = 0.0;
    if (j < M) { qx = (double)pc1[j * 3 + 0]; qy = (double)pc1[j * 3 + 1]; qz = (double)pc1[j * 3 + 2]; }
    double best = __longlong_as_double(0x7ff0000000000000ll);
    int bi = 0;
    for (int i0 = 0; i0 < N; i0 += NN3_TILE) {
        const int n = N - i0 < NN3_TILE ? N - i0 : NN3_TILE;
        __syncthreads();
        for (int e = threadIdx.x; e < n * 3; e += NN3_THREADS) tile[e] = (double)pc0[(size_t)i0 * 3 + e];
        __syncthreads();
        if (j < M) {
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
                const double dx = __dsub_rn(tile[i * 3 + 0], qx), dy = __dsub_rn(tile[i * 3 + 1], qy),
                             dz = __dsub_rn(tile[i * 3 + 2], qz);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (d2 < best) { best = d2; bi = i0 + i; }   // ascending index: the first minimum keeps its place
            }
        }
    }
    bool in = false;
    if (j < M) {
        const double d = __dsqrt_rn(best);
        idx[j] = bi;
        dist[j] = d;
        in = d < thr;
        if (mask) mask[j] = in ? 1 : 0;
    }
    if (count) {
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
    }
}

__global__ void __launch_bounds__(256) transform_kernel(const float *__restrict__ Rt, float *pc, int M)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const double x = (double)pc[j * 3 + 0], y = (double)pc[j * 3 + 1], z = (double)pc[j * 3 + 2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double r0 = (double)Rt[a * 3 + 0], r1 = (double)Rt[a * 3 + 1], r2 = (double)Rt[a * 3 + 2], t = (double)Rt[9 + a];
        const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(r0, x), __dmul_rn(r1, y)), __dmul_rn(r2, z)), t);
        pc[j * 3 + a] = (float)v;
    }
}

}  // namespace

extern "C" int caelo_nn3(caelo_ctx *ctx, const float *pc0, int N, const float *pc1, int M, int64_t *idx, double *dist,
                         double thr, uint8_t *mask, int32_t *count, void *stream)
{
    if (!ctx || !pc0 || !pc1 || !idx || !dist || N <= 0 || M <= 0) return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (count) CAELO_CUDA(ctx, cudaMemsetAsync(count, 0, 4, st));
    { ProfScope ps_(ctx, "nn3_kernel", st);
      nn3_kernel<<<(M + NN3_THREADS - 1) / NN3_THREADS, NN3_THREADS, 0, st>>>(pc0, N, pc1, M, reinterpret_cast<long long *>(idx),
                                                                              dist, thr, mask, count); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}

extern "C" int caelo_transform_points(caelo_ctx *ctx, const float *Rt, float *pc, int M, void *stream)
{
    if (!ctx || !Rt || !pc || M <= 0) return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    { ProfScope ps_(ctx, "transform_kernel", st); transform_kernel<<<(M + 255) / 256, 256, 0, st>>>(Rt, pc, M); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}
