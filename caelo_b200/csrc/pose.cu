// pose.cu — a5: RANSAC4RT / SolveRT (reference Match.py:138-218, 273-283) for sm_100a.
//
// COMPILED WITH -fmad=false: contract K1 (oracle/caelo_oracle.c) is written in plain float64
// +,-,*,/,sqrt and must not be contracted into FMAs, so that gcc and nvcc agree bit-for-bit.
// The float32 scoring contract D1 uses explicit __fmaf_rn where a fused op is specified.
//
//   hyp_score_kernel   one warp per hypothesis: Kabsch of the 4 sampled pairs (the reference's
//                      reflection quirk included), then inlier count over all N pairs with
//                      per-lane counters and a warp-shuffle reduction
//   replay_mask_kernel one CTA per pair: the sequential accept/stop rule replayed over the
//                      per-hypothesis counts, then the inlier mask of the accepted hypothesis
//   kabsch_kernel      one warp per problem: masked refit over all inliers (Match.py:280-282)
#include "common.cuh"
#include "kabsch.cuh"

namespace {

// Kabsch over the points this warp's lanes hold: lane l contributes its own sequential partial
// sums (s0/s1/cnt for the means, then h for H); combination is the K1 butterfly.
struct LanePoint { float p0[3], p1[3]; };

__device__ __forceinline__ bool inlier_d1(const float R[9], const float T[3], const float p0[3],
                                          const float p1[3], float thr)
{
    float e[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float t = __fmul_rn(R[a * 3 + 0], p1[0]);
        t = __fmaf_rn(R[a * 3 + 1], p1[1], t);
        t = __fmaf_rn(R[a * 3 + 2], p1[2], t);
        float q = __fadd_rn(t, T[a]);
        e[a] = __fsub_rn(p0[a], q);
    }
    float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(e[0], e[0]), __fmul_rn(e[1], e[1])), __fmul_rn(e[2], e[2])));
    return d < thr;
}

struct PoseArgs {
    const float *pc0, *pc1;       // [P,N0,3], [P,N,3]
    const long long *pair_idx;    // [P,N] or null
    const int *sample_idx;        // [P,T,4]
    const float *thr;             // [P], or null: thr_scalar for every pair
    float thr_scalar;
    float *thr_out;               // [P] or null: replay writes the threshold of the last round the pair took part in
    const int *best_n_in;         // [P] or null
    const float *skip_if_ok;      // [P,16] or null: pairs with [12] != 0 there are skipped
    int N0, N, T, P;
    int hyp_per_cta;              // hyp_score: trials per CTA (a multiple of the 8 warps)
    int t0, t1;                   // hyp_score: trials [t0,t1) of every pair; replay: trials [0,t1) are scored
    int *more;                    // [P] speculation flag (see caelo_ransac_round): 1 = the pair needs trials >= t1
    int more_mode;                // hyp_score: 1 = only pairs with more[pair]; replay: 1 = first phase (may set more[pair] and
                                  // leave the outputs alone), 2 = second phase (only pairs with more[pair])
    int *counts;                  // [P,T]
    float *rt_hyp;                // [P,T,12]
    float *result;                // [P,16]
    unsigned char *mask;          // [P,N]
};

__device__ __forceinline__ void load_pair(const PoseArgs &a, int pair, int i, float p0[3], float p1[3])
{
    // indices come from the caller (sample_idx, pair_idx): out-of-range values are clamped, never dereferenced
    i = i < 0 ? 0 : (i >= a.N ? a.N - 1 : i);
    long long j = a.pair_idx ? a.pair_idx[(size_t)pair * a.N + i] : i;
    j = j < 0 ? 0 : (j >= a.N0 ? a.N0 - 1 : j);
    const float *q0 = a.pc0 + ((size_t)pair * a.N0 + j) * 3;
    const float *q1 = a.pc1 + ((size_t)pair * a.N + i) * 3;
    p0[0] = q0[0]; p0[1] = q0[1]; p0[2] = q0[2];
    p1[0] = q1[0]; p1[1] = q1[1]; p1[2] = q1[2];
}

// Hypotheses: ONE THREAD per (pair, trial) solves the four-point Kabsch problem (float64 contract K1: with four
// samples the 32-lane butterfly of the refit reduces to (v0 + v2) + (v1 + v3), lanes 4..31 contributing +0.0).
// The float64 Jacobi is a long dependent chain (divisions, square roots): all P*T of them run side by side.
__device__ __forceinline__ double tree4(double v0, double v1, double v2, double v3) { return (v0 + v2) + (v1 + v3); }

__global__ void __launch_bounds__(64) hyp_kabsch_kernel(const PoseArgs a)
{
    const int pair = blockIdx.y;
    const int t = blockIdx.x * 64 + threadIdx.x;
    if (t >= a.T) return;
    if (a.skip_if_ok && a.skip_if_ok[(size_t)pair * 16 + 12] != 0.0f) return;
    const int4 si = *reinterpret_cast<const int4 *>(a.sample_idx + ((size_t)pair * a.T + t) * 4);
    float q0[4][3], q1[4][3];
    load_pair(a, pair, si.x, q0[0], q1[0]); load_pair(a, pair, si.y, q0[1], q1[1]);
    load_pair(a, pair, si.z, q0[2], q1[2]); load_pair(a, pair, si.w, q0[3], q1[3]);
    double m0[3], m1[3];
    for (int c = 0; c < 3; ++c) {
        m0[c] = tree4(0.0 + (double)q0[0][c], 0.0 + (double)q0[1][c], 0.0 + (double)q0[2][c], 0.0 + (double)q0[3][c]) / 4.0;
        m1[c] = tree4(0.0 + (double)q1[0][c], 0.0 + (double)q1[1][c], 0.0 + (double)q1[2][c], 0.0 + (double)q1[3][c]) / 4.0;
    }
    double a0[4][3], a1[4][3];
    for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 3; ++c) { a1[k][c] = (double)q1[k][c] - m1[c]; a0[k][c] = (double)q0[k][c] - m0[c]; }
    double H[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            H[r][c] = tree4(0.0 + a1[0][r] * a0[0][c], 0.0 + a1[1][r] * a0[1][c], 0.0 + a1[2][r] * a0[2][c],
                            0.0 + a1[3][r] * a0[3][c]);
    float R[9], T[3];
    kabsch_from_H(H, m0, m1, R, T);
    float4 *o = reinterpret_cast<float4 *>(a.rt_hyp + ((size_t)pair * a.T + t) * 12);
    o[0] = make_float4(R[0], R[1], R[2], R[3]);
    o[1] = make_float4(R[4], R[5], R[6], R[7]);
    o[2] = make_float4(R[8], T[0], T[1], T[2]);
}

// Scoring: one CTA = one pair x HS_HYP consecutive trials.  The pair's matched points are staged once in shared
// memory (SoA) and the eight warps count inliers, one warp per hypothesis at a time (contract D1).
constexpr int HS_HYP = 8;     // default trials per CTA = one per warp: many small CTAs hide the staging latency best
constexpr int HS_THREADS = 256;

template <bool kStaged>
__global__ void __launch_bounds__(HS_THREADS) hyp_score_kernel(const PoseArgs a)
{
    extern __shared__ float hs_pts[];        // kStaged: [6][N] = x0 y0 z0 x1 y1 z1
    const int pair = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t_base = a.t0 + blockIdx.x * a.hyp_per_cta;
    if (a.skip_if_ok && a.skip_if_ok[(size_t)pair * 16 + 12] != 0.0f) return;
    if (a.more_mode == 1 && !a.more[pair]) return;
    const int N = a.N;
    if (kStaged) {
        for (int i = threadIdx.x; i < N; i += HS_THREADS) {
            float p0[3], p1[3];
            load_pair(a, pair, i, p0, p1);
            hs_pts[i] = p0[0]; hs_pts[N + i] = p0[1]; hs_pts[2 * N + i] = p0[2];
            hs_pts[3 * N + i] = p1[0]; hs_pts[4 * N + i] = p1[1]; hs_pts[5 * N + i] = p1[2];
        }
        __syncthreads();
    }
    const float thr = a.thr ? a.thr[pair] : a.thr_scalar;
    for (int h = warp; h < a.hyp_per_cta && t_base + h < a.t1; h += HS_THREADS / 32) {
        float R[9], T[3];
        {
            const float4 *rt = reinterpret_cast<const float4 *>(a.rt_hyp + ((size_t)pair * a.T + t_base + h) * 12);
            const float4 r0 = __ldg(rt), r1 = __ldg(rt + 1), r2 = __ldg(rt + 2);
            R[0] = r0.x; R[1] = r0.y; R[2] = r0.z; R[3] = r0.w; R[4] = r1.x; R[5] = r1.y; R[6] = r1.z; R[7] = r1.w;
            R[8] = r2.x; T[0] = r2.y; T[1] = r2.z; T[2] = r2.w;
        }
        int cnt = 0;
        for (int i = lane; i < N; i += 32) {
            float p0[3], p1[3];
            if (kStaged) {
                p0[0] = hs_pts[i]; p0[1] = hs_pts[N + i]; p0[2] = hs_pts[2 * N + i];
                p1[0] = hs_pts[3 * N + i]; p1[1] = hs_pts[4 * N + i]; p1[2] = hs_pts[5 * N + i];
            } else {
                load_pair(a, pair, i, p0, p1);
            }
            cnt += inlier_d1(R, T, p0, p1, thr) ? 1 : 0;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) a.counts[(size_t)pair * a.T + t_base + h] = cnt;
    }
}

static_assert(CAELO_MAX_TRIALS <= 2 * 256, "replay_mask_kernel covers two trials per thread of its 256-thread CTA");
__global__ void __launch_bounds__(256) replay_mask_kernel(const PoseArgs a)
{
    __shared__ int s_bt, s_bn, s_ok, s_it, s_more, s_first;
    __shared__ long long s_best;
    __shared__ float s_rt[12];
    const int pair = blockIdx.x;
    if (a.skip_if_ok && a.skip_if_ok[(size_t)pair * 16 + 12] != 0.0f) {
        if (a.more_mode == 1 && threadIdx.x == 0) a.more[pair] = 0;
        return;
    }
    if (a.more_mode == 2 && !a.more[pair]) return;
    // RANSAC4RT's loop (Match.py:181-206) over the pre-scored trials, evaluated in closed form by the whole CTA:
    //   eff_j = cnt_j if cnt_j >= leastInliers else "skipped";  bn_i = max(bn0, eff_0..eff_{i-1}) never decreases, so
    //   the loop `while it < t1 and (it < 100 or (it < 500 and bn < succ))` ends at
    //   it = min(t1, max(100, j*+1), 500) with j* the first trial whose eff reaches succ (or bn0 >= succ: 100);
    //   bn / bt are the maximum over the trials before `it` (first index on ties, only if it beats bn0).
    const int N = a.N;
    int least = (int)(0.2 * (double)N);
    if (least > 100) least = 100;
    const double succ = 0.25 * (double)N;
    const int bn0 = a.best_n_in ? a.best_n_in[pair] : 0;
    const int *cnt = a.counts + (size_t)pair * a.T;
    if (threadIdx.x == 0) { s_first = 0x7fffffff; s_best = -1ll; }
    __syncthreads();
    int eff[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int j = threadIdx.x + u * 256;
        eff[u] = -1;
        if (j < a.t1) {
            const int n = cnt[j];
            if (n >= least) eff[u] = n;
            if (eff[u] >= 0 && !((double)eff[u] < succ)) atomicMin(&s_first, j);
        }
    }
    __syncthreads();
    int it = a.t1;
    {
        int stop = 500;
        if (!((double)bn0 < succ)) stop = 100;
        else if (s_first != 0x7fffffff) stop = s_first + 1 > 100 ? s_first + 1 : 100;
        if (stop < it) it = stop;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int j = threadIdx.x + u * 256;
        if (j < it && eff[u] >= 0) atomicMax(&s_best, ((long long)eff[u] << 10) | (long long)(1023 - j));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int bn = bn0, bt = -1, ok = 0;
        if (s_best >= 0) {
            ok = 1;
            const int n = (int)(s_best >> 10);
            if (n > bn0) { bn = n; bt = 1023 - (int)(s_best & 1023); }
        }
        // first phase: the loop ran out of scored trials while the reference would go on -> the second phase decides
        int more = 0;
        if (a.more_mode == 1) {
            more = (it == a.t1 && it < a.T && ((it < 100) || (it < 500 && (double)bn < succ))) ? 1 : 0;
            a.more[pair] = more;
        }
        s_bt = bt; s_bn = bn; s_ok = ok; s_it = it; s_more = more;
    }
    __syncthreads();
    if (s_more) return;
    const int bt = s_bt;
    if (threadIdx.x < 12) {
        float v = (threadIdx.x == 0 || threadIdx.x == 4 || threadIdx.x == 8) ? 1.0f : 0.0f;  // R_star = I, T_star = 0
        if (bt >= 0) v = a.rt_hyp[((size_t)pair * a.T + bt) * 12 + threadIdx.x];
        s_rt[threadIdx.x] = v;
        a.result[(size_t)pair * 16 + threadIdx.x] = v;
    }
    if (threadIdx.x == 12) a.result[(size_t)pair * 16 + 12] = (float)s_ok;
    if (threadIdx.x == 13) a.result[(size_t)pair * 16 + 13] = (float)s_it;
    if (threadIdx.x == 14) a.result[(size_t)pair * 16 + 14] = (float)s_bn;
    if (threadIdx.x == 15) a.result[(size_t)pair * 16 + 15] = (float)bt;
    __syncthreads();
    float R[9], T[3];
    for (int i = 0; i < 9; ++i) R[i] = s_rt[i];
    for (int i = 0; i < 3; ++i) T[i] = s_rt[9 + i];
    const float thr = a.thr ? a.thr[pair] : a.thr_scalar;
    if (threadIdx.x == 0 && a.thr_out) a.thr_out[pair] = thr;   // last round a pair took part in: the one that gave its model, or the final one
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
        unsigned char m = 0;
        if (bt >= 0) {
            float p0[3], p1[3];
            load_pair(a, pair, i, p0, p1);
            m = inlier_d1(R, T, p0, p1, thr) ? 1 : 0;
        }
        a.mask[(size_t)pair * a.N + i] = m;
    }
}

struct KabschArgs {
    const float *pc0, *pc1;
    const long long *pair_idx;
    const unsigned char *mask;
    const float *skip_if_ok;  // [P,16] or null
    int N0, N, P;
    float *rt;      // [P,12]
    int *credible;  // [P]
};

// One CTA per problem.  The points go through shared memory in chunks of KB_CHUNK: all 256 threads gather a chunk
// (the gathers through pair_idx are two dependent L2 round trips per point: many side by side instead of one after
// the other per lane), then warp 0 runs contract K1 on it: lane l sums points l, l+32, ... in ascending order
// (chunk boundaries are multiples of 32, so the order is the same as over the whole array), xor-butterfly across
// lanes at the end of each of the two passes (means, then H).
constexpr int KB_THREADS = 256;
constexpr int KB_CHUNK = 1024;   // 25 KB of static shared memory (the pipeline's N = 1024 key points fit in one chunk)

__global__ void __launch_bounds__(KB_THREADS) kabsch_kernel(const KabschArgs a)
{
    __shared__ float kb_pts[6][KB_CHUNK];   // x0 y0 z0 x1 y1 z1
    __shared__ unsigned char kb_mask[KB_CHUNK];
    __shared__ double s_mean[6];
    __shared__ int s_cnt;
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.x;
    if (a.skip_if_ok && a.skip_if_ok[(size_t)pair * 16 + 12] != 0.0f) return;
    const int N = a.N;
    auto stage = [&](int c0) {
        __syncthreads();                      // warp 0 is done with the previous chunk
        for (int e = threadIdx.x; e < KB_CHUNK && c0 + e < N; e += KB_THREADS) {
            const int i = c0 + e;
            const unsigned char m = a.mask ? a.mask[(size_t)pair * N + i] : 1;
            kb_mask[e] = m;
            if (m) {
                const long long j = a.pair_idx ? a.pair_idx[(size_t)pair * N + i] : i;
                const float *q0 = a.pc0 + ((size_t)pair * a.N0 + j) * 3;
                const float *q1 = a.pc1 + ((size_t)pair * N + i) * 3;
                kb_pts[0][e] = q0[0]; kb_pts[1][e] = q0[1]; kb_pts[2][e] = q0[2];
                kb_pts[3][e] = q1[0]; kb_pts[4][e] = q1[1]; kb_pts[5][e] = q1[2];
            }
        }
        __syncthreads();
    };
    // ---- pass 1: means ----
    double s[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int c0 = 0; c0 < N; c0 += KB_CHUNK) {
        stage(c0);
        if (threadIdx.x < 32) {
            const int n = N - c0 < KB_CHUNK ? N - c0 : KB_CHUNK;
            for (int e = lane; e < n; e += 32) {
                if (!kb_mask[e]) continue;
                for (int c = 0; c < 3; ++c) { s[c] = s[c] + (double)kb_pts[c][e]; s[3 + c] = s[3 + c] + (double)kb_pts[3 + c][e]; }
                ++cnt;
            }
        }
    }
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        double m[6];
        for (int c = 0; c < 6; ++c) m[c] = warp_tree(s[c]);
        if (lane == 0) {
            s_cnt = cnt;
            for (int c = 0; c < 6; ++c) s_mean[c] = cnt ? m[c] / (double)cnt : 0.0;
        }
    }
    __syncthreads();
    cnt = s_cnt;
    if (cnt == 0) {
        if (threadIdx.x < 12) a.rt[(size_t)pair * 12 + threadIdx.x] = (threadIdx.x == 0 || threadIdx.x == 4 || threadIdx.x == 8) ? 1.0f : 0.0f;
        if (threadIdx.x == 0) a.credible[pair] = 0;
        return;
    }
    double m0[3], m1[3];
    for (int c = 0; c < 3; ++c) { m0[c] = s_mean[c]; m1[c] = s_mean[3 + c]; }
    // ---- pass 2: H ----
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int c0 = 0; c0 < N; c0 += KB_CHUNK) {
        if (N > KB_CHUNK || c0 > 0) stage(c0);      // a single chunk is still in place from pass 1
        if (threadIdx.x < 32) {
            const int n = N - c0 < KB_CHUNK ? N - c0 : KB_CHUNK;
            for (int e = lane; e < n; e += 32) {
                if (!kb_mask[e]) continue;
                double a1[3], a0[3];
                for (int c = 0; c < 3; ++c) { a1[c] = (double)kb_pts[3 + c][e] - m1[c]; a0[c] = (double)kb_pts[c][e] - m0[c]; }
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) h[r * 3 + c] = h[r * 3 + c] + a1[r] * a0[c];
            }
        }
    }
    if (threadIdx.x >= 32) return;
    double H[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) H[r][c] = warp_tree(h[r * 3 + c]);
    float R[9], T[3];
    int cred = kabsch_from_H(H, m0, m1, R, T);
    if (lane < 12) a.rt[(size_t)pair * 12 + lane] = lane < 9 ? R[lane] : T[lane - 9];
    if (lane == 0) a.credible[pair] = cred;
}

// ---- RANSAC sample indices on the device ------------------------------------------------------
// RANSAC4RT draws `np.random.random((4,))` per trial and uses int32(u*N) (Match.py:182-184).  The batched
// pipeline seeds numpy's legacy generator per pair (np.random.seed(pair_id), the harness convention), so the
// whole index stream of a pair is a pure function of the seed: MT19937 init_genrand(seed), then per double
// a = next()>>5, b = next()>>6, u = (a*2^26 + b) / 2^53 (numpy's legacy random_sample).  One CTA per pair:
// thread 0 runs the sequential seeding recurrence, the 624-word twist is done in three data-parallel phases
// (words [0,227) need only old words, [227,454) and [454,624) need new words 227 places back), tempering and
// the conversion are element-wise.
constexpr int MT_N = 624, MT_M = 397;

struct DrawArgs {
    const long long *seeds;   // dev [P]
    int *samples;             // dev [rounds,P,T*4]
    int P, per_round;         // doubles per round = T*4
    int first, count;         // doubles [first, first+count) of every pair's stream are emitted
    int n_points;
};

__device__ __forceinline__ unsigned mt_mix(unsigned hi_word, unsigned lo_word, unsigned far_word)
{
    const unsigned y = (hi_word & 0x80000000u) | (lo_word & 0x7fffffffu);
    return far_word ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__global__ void __launch_bounds__(256) draw_samples_kernel(const DrawArgs a)
{
    __shared__ unsigned key[2][MT_N];
    const int p = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        unsigned s = (unsigned)a.seeds[p];
        for (int i = 0; i < MT_N; ++i) {
            key[0][i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + (unsigned)(i + 1);
        }
    }
    __syncthreads();
    const int last = a.first + a.count;          // doubles
    const int twists = (2 * last + MT_N - 1) / MT_N;
    int cur = 0;
    for (int t = 0; t < twists; ++t) {
        const unsigned *o = key[cur];
        unsigned *n = key[cur ^ 1];
        for (int i = tid; i < MT_N - MT_M; i += 256) n[i] = mt_mix(o[i], o[i + 1], o[i + MT_M]);
        __syncthreads();
        for (int i = MT_N - MT_M + tid; i < 2 * (MT_N - MT_M); i += 256) n[i] = mt_mix(o[i], o[i + 1], n[i - (MT_N - MT_M)]);
        __syncthreads();
        for (int i = 2 * (MT_N - MT_M) + tid; i < MT_N - 1; i += 256) n[i] = mt_mix(o[i], o[i + 1], n[i - (MT_N - MT_M)]);
        if (tid == 0) n[MT_N - 1] = mt_mix(o[MT_N - 1], n[0], n[MT_M - 1]);
        __syncthreads();
        // words [624 t, 624 t + 624) -> doubles [312 t, 312 t + 312)
        for (int j = tid; j < MT_N / 2; j += 256) {
            const int k = t * (MT_N / 2) + j;
            if (k < a.first || k >= last) continue;
            unsigned w[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                unsigned y = n[2 * j + h];
                y ^= y >> 11;
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                y ^= y >> 18;
                w[h] = y;
            }
            const double u = __ddiv_rn(__dadd_rn(__dmul_rn((double)(w[0] >> 5), 67108864.0), (double)(w[1] >> 6)),
                                       9007199254740992.0);
            const int e = k - a.first, r = e / a.per_round, rem = e - r * a.per_round;
            a.samples[((size_t)r * a.P + p) * a.per_round + rem] = (int)__dmul_rn(u, (double)a.n_points);
        }
        cur ^= 1;
        // the next twist overwrites key[cur ^ 1] = this twist's `o`: every thread is past its reads of `o` (barriers above)
    }
}

}  // namespace

int caelo_pose_init(caelo_ctx *ctx)
{
    CAELO_CUDA(ctx, cudaFuncSetAttribute(hyp_score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    return CAELO_OK;
}

// one threshold round for the pairs of `a` (pc0 .. skip_if_ok, thr, result, mask set by the caller)
static int run_round(caelo_ctx *ctx, PoseArgs a, int32_t *counts, cudaStream_t st)
{
    const int P = a.P, T = a.T, N = a.N;
    size_t need = (size_t)P * T * 12 * 4 + (size_t)P * T * 4 + (size_t)P * 4;
    int rc = caelo_reserve(ctx, ctx->pose_ws, need);
    if (rc) return rc;
    a.rt_hyp = reinterpret_cast<float *>(ctx->pose_ws.ptr);
    a.counts = counts ? counts : reinterpret_cast<int *>(a.rt_hyp + (size_t)P * T * 12);
    a.more = reinterpret_cast<int *>(a.rt_hyp + (size_t)P * T * 12) + (size_t)P * T;
    // The reference stops after 100 trials once a model with >= 25 % inliers exists (Match.py:181): score the
    // first 100 trials, replay the accept/stop rule, and only the pairs whose loop would go on get trials
    // 100..T scored and the rule replayed over all of them (same result as scoring everything up front).
    const int T1 = (T > 100 && !counts) ? 100 : T;   // a caller asking for every count gets every trial scored
    const size_t hs_smem = (size_t)N * 24;
    const bool staged = hs_smem <= 160 * 1024;
    {
        const char *e = getenv("CAELO_HS_HYP");   // debug switch for A/B timing
        a.hyp_per_cta = e ? atoi(e) : HS_HYP;
        if (a.hyp_per_cta < 8) a.hyp_per_cta = 8;
    }
    auto score = [&](int n_trials) {
        const dim3 grid((n_trials + a.hyp_per_cta - 1) / a.hyp_per_cta, P);
        ProfScope ps_(ctx, "hyp_score_kernel", st);
        if (staged) hyp_score_kernel<true><<<grid, HS_THREADS, hs_smem, st>>>(a);
        else hyp_score_kernel<false><<<grid, HS_THREADS, 0, st>>>(a);
    };
    a.t0 = 0; a.t1 = T1; a.more_mode = 0;
    { ProfScope ps_(ctx, "hyp_kabsch_kernel", st); hyp_kabsch_kernel<<<dim3((T + 63) / 64, P), 64, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    score(T1);
    CAELO_LAUNCH_CHECK(ctx);
    a.more_mode = T1 < T ? 1 : 0;
    { ProfScope ps_(ctx, "replay_mask_kernel", st); replay_mask_kernel<<<P, 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    if (T1 < T) {
        a.t0 = T1; a.t1 = T; a.more_mode = 1;
        score(T - T1);
        CAELO_LAUNCH_CHECK(ctx);
        a.more_mode = 2;
        { ProfScope ps_(ctx, "replay_mask_kernel", st); replay_mask_kernel<<<P, 256, 0, st>>>(a); }
        CAELO_LAUNCH_CHECK(ctx);
    }
    return CAELO_OK;
}

extern "C" int caelo_ransac_round(caelo_ctx *ctx, const float *pc0, int N0, const float *pc1, int N,
                                  const int64_t *pair_idx, const int32_t *sample_idx, int T,
                                  const float *thr, const int32_t *best_n_in, const float *skip_if_ok, int P,
                                  float *result, uint8_t *inlier_mask, int32_t *counts, void *stream)
{
    if (!ctx || !pc0 || !pc1 || !sample_idx || !thr || !result || !inlier_mask) return CAELO_ERR_ARG;
    if (P <= 0 || N <= 0 || N0 <= 0 || T <= 0 || T > CAELO_MAX_TRIALS) return CAELO_ERR_ARG;
    if (!pair_idx && N0 != N) return CAELO_ERR_ARG;
    PoseArgs a;
    a.pc0 = pc0; a.pc1 = pc1; a.pair_idx = reinterpret_cast<const long long *>(pair_idx);
    a.sample_idx = sample_idx; a.thr = thr; a.thr_scalar = 0.f; a.thr_out = nullptr;
    a.best_n_in = best_n_in; a.skip_if_ok = skip_if_ok;
    a.N0 = N0; a.N = N; a.T = T; a.P = P;
    a.result = result; a.mask = inlier_mask;
    return run_round(ctx, a, counts, (cudaStream_t)stream);
}

static int launch_kabsch(caelo_ctx *ctx, const KabschArgs &a, cudaStream_t st)
{
    { ProfScope ps_(ctx, "kabsch_kernel", st); kabsch_kernel<<<a.P, KB_THREADS, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}

// The whole threshold ladder of RANSAC4RT + the refit of SolveRelativePose for P pairs in one call: round r runs
// only for the pairs that still have no model (their result row is left untouched by the others), the inlier
// mask of the round that produced the model stays in place, and one masked Kabsch at the end refits every pair.
extern "C" int caelo_ransac_ladder(caelo_ctx *ctx, const float *pc0, int N0, const float *pc1, int N,
                                   const int64_t *pair_idx, const int32_t *sample_idx, int T, int rounds,
                                   const float *thr_ladder, int P, float *result, uint8_t *inlier_mask, float *Rt,
                                   float *thr_used, int32_t *credible, void *stream)
{
    if (!ctx || !pc0 || !pc1 || !sample_idx || !thr_ladder || !result || !inlier_mask || !Rt || !thr_used || !credible)
        return CAELO_ERR_ARG;
    if (P <= 0 || N <= 0 || N0 <= 0 || T <= 0 || T > CAELO_MAX_TRIALS || rounds <= 0 || rounds > 8) return CAELO_ERR_ARG;
    if (!pair_idx && N0 != N) return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    PoseArgs a;
    a.pc0 = pc0; a.pc1 = pc1; a.pair_idx = reinterpret_cast<const long long *>(pair_idx);
    a.thr = nullptr; a.thr_out = thr_used; a.best_n_in = nullptr;
    a.N0 = N0; a.N = N; a.T = T; a.P = P;
    a.result = result; a.mask = inlier_mask;
    for (int r = 0; r < rounds; ++r) {
        a.sample_idx = sample_idx + (size_t)r * P * T * 4;
        a.thr_scalar = thr_ladder[r];
        a.skip_if_ok = r == 0 ? nullptr : result;     // in place: rows with isSuccess != 0 are final
        int rc = run_round(ctx, a, nullptr, st);
        if (rc) return rc;
    }
    KabschArgs k;
    k.pc0 = pc0; k.pc1 = pc1; k.pair_idx = a.pair_idx; k.mask = inlier_mask; k.skip_if_ok = nullptr;
    k.N0 = N0; k.N = N; k.P = P; k.rt = Rt; k.credible = credible;
    return launch_kabsch(ctx, k, st);
}

extern "C" int caelo_kabsch(caelo_ctx *ctx, const float *pc0, int N0, const float *pc1, int N,
                            const int64_t *pair_idx, const uint8_t *mask, const float *skip_if_ok, int P,
                            float *Rt, int32_t *credible, void *stream)
{
    if (!ctx || !pc0 || !pc1 || !Rt || !credible || P <= 0 || N <= 0 || N0 <= 0) return CAELO_ERR_ARG;
    if (!pair_idx && N0 != N) return CAELO_ERR_ARG;
    KabschArgs a;
    a.pc0 = pc0; a.pc1 = pc1; a.pair_idx = reinterpret_cast<const long long *>(pair_idx);
    a.mask = mask; a.skip_if_ok = skip_if_ok; a.N0 = N0; a.N = N; a.P = P; a.rt = Rt; a.credible = credible;
    return launch_kabsch(ctx, a, (cudaStream_t)stream);
}

extern "C" int caelo_ransac_draw_samples(caelo_ctx *ctx, const int64_t *seeds, int P, int n_points, int T, int rounds,
                                         int rounds_done, int32_t *samples, void *stream)
{
    if (!ctx || !seeds || !samples || P <= 0 || n_points <= 0 || T <= 0 || T > CAELO_MAX_TRIALS || rounds <= 0 ||
        rounds_done < 0)
        return CAELO_ERR_ARG;
    for (int i = 0; i < P; ++i)
        if (seeds[i] < 0 || seeds[i] > 0xffffffffll) return CAELO_ERR_ARG;  // numpy: "Seed must be between 0 and 2**32 - 1"
    cudaStream_t st = (cudaStream_t)stream;
    int rc = caelo_reserve(ctx, ctx->seed_ws, (size_t)P * 8);
    if (rc) return rc;
    void *h_stage = nullptr;
    cudaEvent_t ev;
    if ((rc = caelo_stage_acquire(ctx, (size_t)P * 8, &h_stage, &ev))) return rc;
    memcpy(h_stage, seeds, (size_t)P * 8);
    CAELO_CUDA(ctx, caelo_stage_copy_async(ctx->seed_ws.ptr, h_stage, (size_t)P * 8, st));
    CAELO_CUDA(ctx, cudaEventRecord(ev, st));
    DrawArgs a;
    a.seeds = reinterpret_cast<const long long *>(ctx->seed_ws.ptr); a.samples = samples; a.P = P;
    a.per_round = T * 4; a.first = rounds_done * T * 4; a.count = rounds * T * 4; a.n_points = n_points;
    { ProfScope ps_(ctx, "draw_samples_kernel", st); draw_samples_kernel<<<P, 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}
