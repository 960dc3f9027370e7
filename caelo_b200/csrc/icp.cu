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
// Those two are the per-iteration primitives (the reference's own loop driven from the host, kept as the fallback).
//
// caelo_icp_batch runs WHOLE ICPs for a batch of frame pairs on the device, no host round trip per iteration:
//   icp_grid_*        a uniform grid over every pair's PC0 (cell = a hair more than the initial inlier threshold, the
//                     threshold only ever decays): cells live in a per-pair hash table, points are counting-sorted
//                     by cell.  A PC1 point's nearest PC0 point within the threshold, if there is one, is in its 27
//                     neighbouring cells — so the inlier pairs are exactly the brute-force ones (contract N1
//                     arithmetic, ties -> lowest index), at ~1 % of the distance evaluations
//   icp_nn_kernel     per PC1 point: apply the previous iteration's [R|T] (contract U1), search the 27 cells
//   icp_solve_kernel  per pair: SolveRT on the inlier pairs (contract K1, same lane order as kabsch_kernel), then the
//                     reference's loop control (MyICP.py:40-66) by one thread: < 100 inliers -> failure, Euler-angle /
//                     translation norms, convergence after minIterTimes, threshold decay; [R|T] and the inlier count of
//                     every iteration are recorded so that the host can accumulate R*/T* with the reference's own
//                     numpy expressions and re-check every decision the device took
#include "common.cuh"
#include "kabsch.cuh"

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


__global__ void icp_init_kernel(double *thr, int B, double thr0)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) thr[b] = thr0;
}

__global__ void icp_state_kernel(const double *thr, const int *success, const int *iters, const int *last_n, int B, double *state)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    state[b * 4 + 0] = (double)success[b];
    state[b * 4 + 1] = (double)iters[b];
    state[b * 4 + 2] = (double)last_n[b];
    state[b * 4 + 3] = thr[b];
}

// ---- batched ICP -------------------------------------------------------------------------------------------
constexpr unsigned long long ICP_EMPTY = ~0ull;
constexpr double RAD2DEG = 180.0 / 3.14159265358979323846;      // Transformations.py: RADIAN2DEGREE = 180 / pi

struct IcpArgs {
    const float *pc0;            // [S0,3]
    float *pc1;                  // [S1,3] work copy, updated in place
    const int *off0, *off1;      // [B+1] rows
    const int *tab_off;          // [B+1] slots
    int B, S0, S1;
    double inv_cell;             // 1 / cell size
    unsigned long long *keys;    // [slots]
    int4 *meta;                  // [slots] start (absolute row of `sorted`), count, fill, -
    int *pt_slot;                // [S0]
    float4 *sorted;              // [S0] x y z, local index bits
    int *nn;                     // [S1] local PC0 index of the nearest point inside the threshold (-1: none)
    // per-pair state
    double *thr;                 // [B]
    int *done, *apply_it, *iters, *success, *last_n;
    float *rt;                   // [B,12] the current iteration's [R|T]
    float *hist;                 // [B,max_iter,12]
    int *hist_n;                 // [B,max_iter]
    int it, max_iter, min_iter, min_inliers;
    double decay, small_shift, ep;
};

__device__ __forceinline__ int icp_pair_of(const int *off, int B, int row)
{
    int lo = 0, hi = B;          // off[lo] <= row < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= row) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ unsigned long long icp_key(long long ix, long long iy, long long iz)
{
    return ((unsigned long long)(ix + (1 << 20)) << 42) | ((unsigned long long)(iy + (1 << 20)) << 21) |
           (unsigned long long)(iz + (1 << 20));
}

__device__ __forceinline__ unsigned icp_hash(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

__global__ void __launch_bounds__(256) icp_grid_count_kernel(const IcpArgs a)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.S0) return;
    const int b = icp_pair_of(a.off0, a.B, i);
    const long long ix = (long long)floor((double)a.pc0[i * 3 + 0] * a.inv_cell), iy = (long long)floor((double)a.pc0[i * 3 + 1] * a.inv_cell),
                    iz = (long long)floor((double)a.pc0[i * 3 + 2] * a.inv_cell);
    const unsigned long long key = icp_key(ix, iy, iz);
    const int t0 = a.tab_off[b];
    const unsigned m = (unsigned)(a.tab_off[b + 1] - t0) - 1u;
    unsigned s = icp_hash(key) & m;
    while (true) {
        const unsigned long long prev = atomicCAS(a.keys + t0 + s, ICP_EMPTY, key);
        if (prev == ICP_EMPTY || prev == key) break;
        s = (s + 1) & m;
    }
    atomicAdd(&a.meta[t0 + s].y, 1);
    a.pt_slot[i] = t0 + (int)s;
}

__global__ void __launch_bounds__(256) icp_grid_start_kernel(const IcpArgs a, int n_slots, int *pair_fill)
{
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= n_slots || a.keys[s] == ICP_EMPTY) return;
    const int b = icp_pair_of(a.tab_off, a.B, s);
    a.meta[s].x = a.off0[b] + atomicAdd(pair_fill + b, a.meta[s].y);
}

__global__ void __launch_bounds__(256) icp_grid_fill_kernel(const IcpArgs a)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.S0) return;
    const int b = icp_pair_of(a.off0, a.B, i);
    int4 *m = a.meta + a.pt_slot[i];
    const int pos = m->x + atomicAdd(&m->z, 1);
    a.sorted[pos] = make_float4(a.pc0[i * 3 + 0], a.pc0[i * 3 + 1], a.pc0[i * 3 + 2], __int_as_float(i - a.off0[b]));
}

__global__ void __launch_bounds__(128) icp_nn_kernel(const IcpArgs a)
{
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= a.S1) return;
    const int b = icp_pair_of(a.off1, a.B, j);
    float x = a.pc1[j * 3 + 0], y = a.pc1[j * 3 + 1], z = a.pc1[j * 3 + 2];
    if (a.it > 0 && a.apply_it[b] == a.it - 1) {        // PC1 = (R PC1^T + T)^T of the previous iteration (contract U1)
        const float *Rt = a.rt + b * 12;
        const double dx = (double)x, dy = (double)y, dz = (double)z;
        float o[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            o[r] = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn((double)Rt[r * 3], dx), __dmul_rn((double)Rt[r * 3 + 1], dy)),
                                              __dmul_rn((double)Rt[r * 3 + 2], dz)), (double)Rt[9 + r]);
        x = o[0]; y = o[1]; z = o[2];
        a.pc1[j * 3 + 0] = x; a.pc1[j * 3 + 1] = y; a.pc1[j * 3 + 2] = z;
    }
    if (a.done[b]) return;
    const double qx = (double)x, qy = (double)y, qz = (double)z;
    const long long cx = (long long)floor(qx * a.inv_cell), cy = (long long)floor(qy * a.inv_cell), cz = (long long)floor(qz * a.inv_cell);
    const int t0 = a.tab_off[b];
    const unsigned m = (unsigned)(a.tab_off[b + 1] - t0) - 1u;
    double best = __longlong_as_double(0x7ff0000000000000ll);
    int bi = -1;
    // only the cells the threshold ball can reach: the neighbour below / above on an axis is needed iff the query is
    // closer than thr to that face of its own cell (thr <= cell; a point AT distance thr is not an inlier)
    const double thr = a.thr[b], tc = thr * a.inv_cell + 1e-9;
    const double fx = qx * a.inv_cell - (double)cx, fy = qy * a.inv_cell - (double)cy, fz = qz * a.inv_cell - (double)cz;
    const int x0 = fx <= tc ? -1 : 0, x1 = 1.0 - fx <= tc ? 1 : 0, y0 = fy <= tc ? -1 : 0, y1 = 1.0 - fy <= tc ? 1 : 0,
              z0 = fz <= tc ? -1 : 0, z1 = 1.0 - fz <= tc ? 1 : 0;
    for (int ex = x0; ex <= x1; ++ex)
    for (int ey = y0; ey <= y1; ++ey)
    for (int ez = z0; ez <= z1; ++ez) {
        const unsigned long long key = icp_key(cx + ex, cy + ey, cz + ez);
        unsigned s = icp_hash(key) & m;
        int start = 0, count = 0;
        while (true) {
            const unsigned long long k = a.keys[t0 + s];
            if (k == key) { const int4 mm = a.meta[t0 + s]; start = mm.x; count = mm.y; break; }
            if (k == ICP_EMPTY) break;
            s = (s + 1) & m;
        }
        for (int c = 0; c < count; ++c) {
            const float4 p = __ldg(a.sorted + start + c);
            const double dx = __dsub_rn((double)p.x, qx), dy = __dsub_rn((double)p.y, qy), dz = __dsub_rn((double)p.z, qz);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            const int idx = __float_as_int(p.w);
            if (d2 < best || (d2 == best && idx < bi)) { best = d2; bi = idx; }   // exact ties -> lowest PC0 index
        }
    }
    a.nn[j] = (bi >= 0 && __dsqrt_rn(best) < thr) ? bi : -1;
}

// SolveRT on the inlier pairs of one ICP: ONE WARP per pair, in the lane order of contract K1 (lane l sums the inliers among
// the PC1 rows l, l+32, ... in ascending order, xor-butterfly at the end — the same sums as kabsch_kernel in pose.cu).  The
// float64 add chains are short (N/32 steps); what costs is the latency of the dependent gathers nn -> pc0, so every lane
// keeps IS_UNROLL rows in flight before it adds them up in order.  Then the loop control.
constexpr int IS_THREADS = 32;
constexpr int IS_UNROLL = 8;

__global__ void __launch_bounds__(IS_THREADS) icp_solve_kernel(const IcpArgs a)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    if (a.done[b]) return;
    const int r0 = a.off1[b], N = a.off1[b + 1] - r0, p0 = a.off0[b];
    double s[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int e0 = lane; e0 < N; e0 += 32 * IS_UNROLL) {
        int nn[IS_UNROLL];
        float q[IS_UNROLL][6];
#pragma unroll
        for (int k = 0; k < IS_UNROLL; ++k) { const int e = e0 + 32 * k; nn[k] = e < N ? a.nn[r0 + e] : -1; }
#pragma unroll
        for (int k = 0; k < IS_UNROLL; ++k)
            if (nn[k] >= 0) {
                const float *q0 = a.pc0 + (size_t)(p0 + nn[k]) * 3, *q1 = a.pc1 + (size_t)(r0 + e0 + 32 * k) * 3;
                q[k][0] = q0[0]; q[k][1] = q0[1]; q[k][2] = q0[2]; q[k][3] = q1[0]; q[k][4] = q1[1]; q[k][5] = q1[2];
            }
#pragma unroll
        for (int k = 0; k < IS_UNROLL; ++k)
            if (nn[k] >= 0) {
#pragma unroll
                for (int c = 0; c < 6; ++c) s[c] = s[c] + (double)q[k][c];
                ++cnt;
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    double m0[3], m1[3];
    for (int c = 0; c < 3; ++c) {
        const double t0 = warp_tree(s[c]), t1 = warp_tree(s[3 + c]);
        m0[c] = __shfl_sync(0xffffffffu, cnt ? t0 / (double)cnt : 0.0, 0);
        m1[c] = __shfl_sync(0xffffffffu, cnt ? t1 / (double)cnt : 0.0, 0);
    }
    if (cnt < a.min_inliers) {                        // MyICP.py:40-42: return R_star, T_star, False
        if (lane == 0) {
            a.hist_n[b * a.max_iter + a.it] = cnt;
            a.last_n[b] = cnt; a.iters[b] = a.it + 1; a.success[b] = 0; a.done[b] = 1;
        }
        return;
    }
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int e0 = lane; e0 < N; e0 += 32 * IS_UNROLL) {
        int nn[IS_UNROLL];
        float q[IS_UNROLL][6];
#pragma unroll
        for (int k = 0; k < IS_UNROLL; ++k) { const int e = e0 + 32 * k; nn[k] = e < N ? a.nn[r0 + e] : -1; }
#pragma unroll
        for (int k = 0; k < IS_UNROLL; ++k)
            if (nn[k] >= 0) {
                const float *q0 = a.pc0 + (size_t)(p0 + nn[k]) * 3, *q1 = a.pc1 + (size_t)(r0 + e0 + 32 * k) * 3;
                q[k][0] = q0[0]; q[k][1] = q0[1]; q[k][2] = q0[2]; q[k][3] = q1[0]; q[k][4] = q1[1]; q[k][5] = q1[2];
            }
#pragma unroll
        for (int k = 0; k < IS_UNROLL; ++k)
            if (nn[k] >= 0) {
                double a1[3], a0[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { a1[c] = (double)q[k][3 + c] - m1[c]; a0[c] = (double)q[k][c] - m0[c]; }
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) h[r * 3 + c] = h[r * 3 + c] + a1[r] * a0[c];
            }
    }
    double H[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) H[r][c] = warp_tree(h[r * 3 + c]);
    float R[9], T[3];
    kabsch_from_H(H, m0, m1, R, T);
    if (lane < 12) {
        const float v = lane < 9 ? R[lane] : T[lane - 9];
        a.rt[b * 12 + lane] = v;
        a.hist[((size_t)b * a.max_iter + a.it) * 12 + lane] = v;
    }
    if (lane == 0) {
        a.hist_n[b * a.max_iter + a.it] = cnt;
        a.last_n[b] = cnt;
        a.apply_it[b] = a.it;
        a.iters[b] = a.it + 1;
        // MyICP.py:56-66 (RotateMat2EulerAngle_XYZ in degrees; LA.norm(T) of the float32 column is a float32 norm)
        const double e0 = atan2((double)R[7], (double)R[8]) * RAD2DEG;
        const double e1 = atan2(-(double)R[6], sqrt((double)R[7] * (double)R[7] + (double)R[8] * (double)R[8])) * RAD2DEG;
        const double e2 = atan2((double)R[3], (double)R[0]) * RAD2DEG;
        const double nE = sqrt((e0 * e0 + e1 * e1) + e2 * e2);
        const double nT = (double)__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], T[0]), __fmul_rn(T[1], T[1])), __fmul_rn(T[2], T[2])));
        bool stop = false;
        if (a.it >= a.min_iter && nE < a.ep && nT < a.ep) stop = true;
        if (!stop && nE < a.small_shift && nT < a.small_shift) a.thr[b] = a.thr[b] * a.decay;
        if (stop || a.it + 1 == a.max_iter) { a.success[b] = 1; a.done[b] = 1; }
    }
}

}  // namespace

extern "C" int caelo_nn3(caelo_ctx *ctx, const float *pc0, int N, const float *pc1, int M, int64_t *idx, double *dist,
                         double thr, uint8_t *mask, int32_t *count, void *stream)
{
    if (!ctx || !pc0 || !pc1 || !idx || !dist || N <= 0 || M <= 0) return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (count) CAELO_CUDA(ctx, caelo_fill_async(count, 0, 4, st));
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

extern "C" int caelo_icp_batch(caelo_ctx *ctx, const float *pc0, const int64_t *off0, float *pc1, const int64_t *off1, int B,
                               double thr0, double decay, double small_shift, double ep, int max_iter, int min_iter,
                               int min_inliers, float *hist, int32_t *hist_n, double *state, void *stream)
{
    if (!ctx || !pc0 || !off0 || !pc1 || !off1 || !hist || !hist_n || !state || B <= 0 || max_iter <= 0 || !(thr0 > 0.0))
        return CAELO_ERR_ARG;
    for (int b = 0; b < B; ++b)
        if (off0[b + 1] <= off0[b] || off1[b + 1] <= off1[b]) return CAELO_ERR_ARG;
    if (off0[B] > 0x7fffffffll / 4 || off1[B] > 0x7fffffffll / 4) return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int S0 = (int)off0[B], S1 = (int)off1[B];
    // host-side layout: per-pair hash tables of 2^ceil(log2(2 N0)) slots
    const size_t n_int = 3 * (size_t)(B + 1);
    void *h_stage = nullptr;
    cudaEvent_t ev;
    int rc = caelo_stage_acquire(ctx, n_int * 4, &h_stage, &ev);
    if (rc) return rc;
    int *h = static_cast<int *>(h_stage);
    long long slots = 0;
    for (int b = 0; b <= B; ++b) {
        h[b] = (int)off0[b];
        h[(B + 1) + b] = (int)off1[b];
        h[2 * (B + 1) + b] = (int)slots;
        if (b < B) {
            long long n = off0[b + 1] - off0[b], s = 64;
            while (s < 2 * n) s <<= 1;
            slots += s;
        }
    }
    if (slots > 0x7fffffffll) return CAELO_ERR_ARG;
    // device scratch: [offsets][keys][meta][pt_slot][sorted][nn][per-pair state]
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_off = 0, o_keys = al(n_int * 4), o_meta = o_keys + al((size_t)slots * 8), o_slot = o_meta + al((size_t)slots * 16),
                 o_sorted = o_slot + al((size_t)S0 * 4), o_nn = o_sorted + al((size_t)S0 * 16), o_state = o_nn + al((size_t)S1 * 4),
                 o_end = o_state + al((size_t)B * (8 + 6 * 4 + 48));
    rc = caelo_reserve(ctx, ctx->icp_ws, o_end);
    if (rc) return rc;
    unsigned char *ws = static_cast<unsigned char *>(ctx->icp_ws.ptr);
    CAELO_CUDA(ctx, caelo_stage_copy_async(ws + o_off, h_stage, n_int * 4, st));
    CAELO_CUDA(ctx, cudaEventRecord(ev, st));
    IcpArgs a;
    a.pc0 = pc0; a.pc1 = pc1; a.B = B; a.S0 = S0; a.S1 = S1;
    a.off0 = reinterpret_cast<int *>(ws + o_off); a.off1 = a.off0 + (B + 1); a.tab_off = a.off0 + 2 * (B + 1);
    a.inv_cell = 1.0 / (thr0 * (1.0 + 1e-6));
    a.keys = reinterpret_cast<unsigned long long *>(ws + o_keys); a.meta = reinterpret_cast<int4 *>(ws + o_meta);
    a.pt_slot = reinterpret_cast<int *>(ws + o_slot); a.sorted = reinterpret_cast<float4 *>(ws + o_sorted);
    a.nn = reinterpret_cast<int *>(ws + o_nn);
    unsigned char *sp = ws + o_state;
    a.thr = reinterpret_cast<double *>(sp); sp += (size_t)B * 8;
    a.done = reinterpret_cast<int *>(sp); a.apply_it = a.done + B; a.iters = a.done + 2 * B; a.success = a.done + 3 * B;
    a.last_n = a.done + 4 * B;
    int *pair_fill = a.done + 5 * B;
    sp += (size_t)B * 6 * 4;
    a.rt = reinterpret_cast<float *>(sp);
    a.hist = hist; a.hist_n = hist_n; a.max_iter = max_iter; a.min_iter = min_iter; a.min_inliers = min_inliers;
    a.decay = decay; a.small_shift = small_shift; a.ep = ep; a.it = 0;
    CAELO_CUDA(ctx, caelo_fill_async(a.keys, 0xFF, (size_t)slots * 8, st));
    CAELO_CUDA(ctx, caelo_fill_async(a.meta, 0, (size_t)slots * 16, st));
    CAELO_CUDA(ctx, caelo_fill_async(a.done, 0, (size_t)B * 6 * 4, st));
    CAELO_CUDA(ctx, caelo_fill_async(a.apply_it, 0xFF, (size_t)B * 4, st));
    CAELO_CUDA(ctx, caelo_fill_async(hist_n, 0, (size_t)B * max_iter * 4, st));
    CAELO_CUDA(ctx, caelo_fill_async(hist, 0, (size_t)B * max_iter * 48, st));
    { ProfScope ps_(ctx, "icp_init_kernel", st); icp_init_kernel<<<(B + 255) / 256, 256, 0, st>>>(a.thr, B, thr0); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "icp_grid_count_kernel", st); icp_grid_count_kernel<<<(S0 + 255) / 256, 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "icp_grid_start_kernel", st);
      icp_grid_start_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(a, (int)slots, pair_fill); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "icp_grid_fill_kernel", st); icp_grid_fill_kernel<<<(S0 + 255) / 256, 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    for (int it = 0; it < max_iter; ++it) {
        a.it = it;
        { ProfScope ps_(ctx, "icp_nn_kernel", st); icp_nn_kernel<<<(S1 + 127) / 128, 128, 0, st>>>(a); }
        CAELO_LAUNCH_CHECK(ctx);
        { ProfScope ps_(ctx, "icp_solve_kernel", st); icp_solve_kernel<<<B, IS_THREADS, 0, st>>>(a); }
        CAELO_LAUNCH_CHECK(ctx);
    }
    { ProfScope ps_(ctx, "icp_state_kernel", st);
      icp_state_kernel<<<(B + 255) / 256, 256, 0, st>>>(a.thr, a.success, a.iters, a.last_n, B, state); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}
