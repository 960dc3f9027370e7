// match.cu — a4: descriptor nearest-neighbour match (reference Match.py:257-258:
// cdist(Codes0, Codes1,'euclidean') in float64, then argmin(axis=0), ties -> lowest row).
//
// Index-exact in two steps: a fast approximate pass that keeps, per column, the best and second-best d^2 and
// decides every column whose runner-up is outside the pass's error margin, then an exact re-scan of the rest.
//   desc_prep_kernel one pass per descriptor set: squared norms + the split-fp16 operands (a = hi + lo) written
//                    once, tile by tile, in the layout the tensor core reads;
//   nn_tc_kernel     (D <= 128) tcgen05: d^2 = |a|^2 + |b|^2 - 2 a.b with the cross term as a split-fp16
//                    GEMM (hi.hi + lo.hi + hi.lo, fp32 accumulators in TMEM); 128 frame-1 descriptors per CTA
//                    are the MMA rows (= TMEM lanes = threads), frame-0 descriptors stream through as
//                    128-column tiles fetched by bulk copies (cp.async.bulk, mbarrier complete_tx; producer
//                    thread / MMA issuer / four scan warps), so each thread scans its own accumulator row for
//                    the running (min, second min, argmin) with no cross-thread traffic.  Margin: |error| is assumed
//                    <= 2^-13 |a|max |b| (about 1000x the analytic split-fp16 bound 3*2^-22 |a||b|).
//   nn_tile_kernel   (any D) float32 direct-difference distances on 64x64 tiles held in shared memory;
//                    per column the best and second-best approximate d^2 (+ best row).  The
//                    float32 sum of D non-negative terms is within (D+4)*2^-24 relative of the
//                    true value, so a column whose runner-up is outside that margin is decided.
//   nn_exact_kernel  one warp per column; undecided columns are re-scanned: a float32 pass keeps the rows
//                    inside the margin, those are evaluated with the reference's own arithmetic
//                    (contract M1: float64 sequential sum, sqrt, lowest row wins).
#include "common.cuh"
#include "umma.cuh"

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
    const float *n1;       // [P,M] |c1 row|^2 and
    const float *nmax0;    // [P] max |c0 row|^2: absolute margins of the tensor-core pass (null: relative float32 margin)
    float margin;          // E = margin * |a|max |b_j| per d^2 value of the tensor-core pass
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


// ---- tensor-core pass ---------------------------------------------------------------------------
constexpr int NT_B = 128;                 // rows (frame-1) per CTA and columns (frame-0) per tile
constexpr int NT_THREADS = 192;           // warps 0-3: scan accumulators; warp 4: MMA issuer; warp 5: bulk-copy producer

struct NNTcArgs {
    const unsigned char *ops0, *ops1;   // split-fp16 operand tiles of frame 0 / frame 1 (desc_prep_kernel)
    const float *n0p;                   // [P, tiles0*128] squared norms of frame 0, +inf in the padding rows
    const float *n1;                    // [P, M]
    int N, M, Kp, tiles0, tiles1;       // Kp = D rounded up to 16
    float *best_d, *second_d;
    int *best_i;
};

// One pass over a descriptor set: squared norms (float64 sums rounded once; the pair's maximum by atomicMax) and
// the split-fp16 MMA operands, written once in the layout the tensor core reads: per 128-row tile one contiguous
// block [hi, lo][chunk of 8 k][row][16 B] (canonical K-major no-swizzle: SBO = 128 B, LBO = 128*16 B), so that a
// tile reaches shared memory with ONE bulk copy and no further conversion.  Rows beyond the set and k >= D are zero.
__global__ void __launch_bounds__(128) desc_prep_kernel(const float *c, int rows_per_pair, int D, int Kp, int tiles,
                                                        unsigned char *ops, float *norm, float *norm_padded, float *pair_max)
{
    const int pair = blockIdx.y, tile = blockIdx.x, r = threadIdx.x;
    const int row = tile * NT_B + r;
    const bool in = row < rows_per_pair;
    const float *p = c + ((size_t)pair * rows_per_pair + row) * D;
    unsigned char *blk = ops + ((size_t)pair * tiles + tile) * ((size_t)Kp * NT_B * 4);
    double acc = 0.0;
    for (int ch = 0; ch < Kp / 8; ++ch) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (in && ch * 8 + k < D) ? __ldg(p + ch * 8 + k) : 0.0f;
        __half2 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::split_f16x2(v[2 * k], v[2 * k + 1], h[k], l[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += (double)v[k] * (double)v[k];
        *reinterpret_cast<uint4 *>(blk + (size_t)ch * (NT_B * 16) + r * 16) = *reinterpret_cast<uint4 *>(h);
        *reinterpret_cast<uint4 *>(blk + (size_t)Kp * NT_B * 2 + (size_t)ch * (NT_B * 16) + r * 16) = *reinterpret_cast<uint4 *>(l);
    }
    const float nv = (float)acc;
    if (in) {
        if (norm) norm[(size_t)pair * rows_per_pair + row] = nv;
        if (pair_max) atomicMax(reinterpret_cast<int *>(pair_max) + pair, __float_as_int(nv));  // nv >= 0: int order = float order
    }
    if (norm_padded) norm_padded[((size_t)pair * tiles + tile) * NT_B + r] = in ? nv : __int_as_float(0x7f800000);
}

__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     umma::smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(umma::smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(mbar)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(NT_THREADS) nn_tc_kernel(const NNTcArgs a)
{
    extern __shared__ __align__(128) unsigned char nsm[];
    const int opb = a.Kp * NT_B * 2;                 // bytes of one operand half (hi or lo); a tile block is 2*opb
    unsigned char *A_blk = nsm;                      // [hi, lo][opb]
    unsigned char *B_base = nsm + 2 * opb;           // [buf 2][hi, lo][opb]
    float *n0s = reinterpret_cast<float *>(nsm + 6 * opb);   // [2][NT_B]
    uint64_t *full = reinterpret_cast<uint64_t *>(n0s + 2 * NT_B), *tfull = full + 2, *tempty = full + 4, *afull = full + 6;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(full + 7);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int pair = blockIdx.y, j0 = blockIdx.x * NT_B;
    const float INF = __int_as_float(0x7f800000);

    if (warp == 4) umma::tmem_alloc(tmem_slot, 2 * NT_B);
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&full[b], 1);             // producer's expect_tx arrival + the bytes of the bulk copies
            umma::mbar_init(&tfull[b], 1);            // MMAs of the tile complete (tcgen05.commit)
            umma::mbar_init(&tempty[b], 128);         // accumulators + norms of the tile consumed by the four scan warps
        }
        umma::mbar_init(afull, 1);
        umma::fence_mbar_init();
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = *tmem_slot;
    const int ntiles = a.tiles0;

    if (warp == 5) {
        // ===== producer: one elected thread streams the operand tiles with bulk copies =====
        if (umma::elect_one()) {
            mbar_expect_tx(afull, 2 * opb);
            bulk_g2s(A_blk, a.ops1 + ((size_t)pair * a.tiles1 + blockIdx.x) * (size_t)(2 * opb), 2 * opb, afull);
            for (int t = 0; t < ntiles; ++t) {
                const int b = t & 1, k = t >> 1;
                if (k >= 1) {   // MMA(t-2) has read the operand buffer, the scan warps have read its norms
                    umma::mbar_wait(&tfull[b], (uint32_t)((k - 1) & 1));
                    umma::mbar_wait(&tempty[b], (uint32_t)((k - 1) & 1));
                }
                mbar_expect_tx(&full[b], 2 * opb + NT_B * 4);
                bulk_g2s(B_base + b * 2 * opb, a.ops0 + ((size_t)pair * a.tiles0 + t) * (size_t)(2 * opb), 2 * opb, &full[b]);
                bulk_g2s(n0s + b * NT_B, a.n0p + ((size_t)pair * a.tiles0 + t) * NT_B, NT_B * 4, &full[b]);
            }
        }
        __syncwarp();
    } else if (warp == 4) {
        const uint32_t idesc = umma::idesc_f16_f32(NT_B, NT_B);
        const uint32_t sA_hi = umma::smem_u32(A_blk), sA_lo = sA_hi + opb, sB = umma::smem_u32(B_base);
        umma::mbar_wait(afull, 0);
        for (int t = 0; t < ntiles; ++t) {
            const int b = t & 1, k = t >> 1;
            umma::mbar_wait(&full[b], (uint32_t)(k & 1));
            if (k >= 1) umma::mbar_wait(&tempty[b], (uint32_t)((k - 1) & 1));
            umma::fence_after_thread_sync();
            if (umma::elect_one()) {
                const uint32_t d = tbase + b * NT_B;
                const uint32_t bh = sB + b * 2 * opb, bl = bh + opb;
                for (int j = 0; j < a.Kp / 16; ++j) {
                    const uint32_t off = j * 2 * (NT_B * 16);
                    const uint64_t ah = umma::smem_desc(sA_hi + off, NT_B * 16, 128), al = umma::smem_desc(sA_lo + off, NT_B * 16, 128);
                    const uint64_t wh = umma::smem_desc(bh + off, NT_B * 16, 128), wl = umma::smem_desc(bl + off, NT_B * 16, 128);
                    umma::mma_f16(d, ah, wh, idesc, j ? 1u : 0u);
                    umma::mma_f16(d, al, wh, idesc, 1u);
                    umma::mma_f16(d, ah, wl, idesc, 1u);
                }
                umma::commit(&tfull[b]);
            }
            __syncwarp();
        }
    } else {
        // ===== scan warps: thread = one frame-1 descriptor (an accumulator row), running (min, second min, argmin) =====
        float best = INF, second = INF;
        int bi = 0x7fffffff;
        for (int t = 0; t < ntiles; ++t) {
            const int b = t & 1;
            umma::mbar_wait(&full[b], (uint32_t)((t >> 1) & 1));    // the tile's norms have landed (bulk copy observed by this thread)
            umma::mbar_wait(&tfull[b], (uint32_t)((t >> 1) & 1));   // its MMAs are complete
            umma::fence_after_thread_sync();
#pragma unroll 1
            for (int cc = 0; cc < NT_B; cc += 32) {
                uint32_t v[32];
                umma::tmem_ld_x32(tbase + ((uint32_t)(32 * warp) << 16) + b * NT_B + cc, v);
                umma::tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const float key = fmaf(-2.0f, __uint_as_float(v[k]), n0s[b * NT_B + cc + k]);  // |a_i|^2 - 2 a_i.b_j
                    second = fminf(second, fmaxf(key, best));
                    if (key < best) bi = t * NT_B + cc + k;   // ascending i: the first minimum keeps its index
                    best = fminf(best, key);
                }
            }
            umma::fence_before_thread_sync();
            umma::mbar_arrive(&tempty[b]);
        }
        const int j = j0 + tid;
        if (j < a.M) {
            const size_t o = (size_t)pair * a.M + j;
            const float n1 = a.n1[o];
            a.best_d[o] = fmaxf(best + n1, 0.0f);
            a.second_d[o] = fmaxf(second + n1, 0.0f);
            a.best_i[o] = bi;
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    if (warp == 4) umma::tmem_dealloc(tbase, 2 * NT_B);
}

// decided columns get their index; the others are appended to a list for the exact re-scan
// Error margin of one d^2 value of the tensor-core pass: a term proportional to |a|max |b_j| (split-fp16 products, fp32
// accumulation in the tensor core) plus an absolute term for the fp16 lo parts that fall into the subnormal range
// (|lo| <= 2^-12 |x| is below 2^-14 for every |x| < 1/4: absolute error <= 2^-25 per element, so <= 2^-25 sqrt(D) (|a| + |b|)
// on the dot product).  Measured (tools/nn_margin.py, profiles/r2_nn_margin.txt): 2^-18.6 |a||b| on tanh-distributed
// descriptors, 2^-15.5 |a||b| on descriptors of magnitude 1e-3 — which the absolute term covers 10x over.
__device__ __forceinline__ float nn_margin_E(const NNArgs &a, int pair, long long w)
{
    if (!a.nmax0) return 0.0f;
    const float na = sqrtf(a.nmax0[pair]), nb = sqrtf(a.n1[w]);
    return a.margin * na * nb + 1.1920929e-7f * sqrtf((float)a.D) * (na + nb) + 1e-30f;
}

__global__ void __launch_bounds__(256) nn_decide_kernel(const NNArgs a, int P, int *list, int *nlist)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (long long)P * a.M) return;
    const int pair = (int)(w / a.M);
    const float b = a.best_d[w], s = a.second_d[w];
    // relative error of each float32 d^2 <= (D+4)*2^-24; undecided if the intervals can overlap.  After the
    // tensor-core pass the error is absolute: E = 2^-13 |a|max |b_j| per value (see the header).
    const float eps = (float)(a.D + 4) * 5.9604645e-8f;
    const float E = nn_margin_E(a, pair, w);
    const bool decided = s > (b + 2.0f * E) * (1.0f + 4.0f * eps) + 1e-30f;
    if (decided) a.out[w] = a.best_i[w];
    else list[atomicAdd(nlist, 1)] = (int)w;
}

// one CTA per undecided column (grid-stride over the list): every thread takes rows tid, tid+256, ...; a float32
// pass keeps the rows inside the margin (any summation order is within (D+4)*2^-24 of the true value), those
// are evaluated with the reference's own arithmetic (contract M1: float64 sequential sum, sqrt), and the CTA
// reduces to the smallest (distance, row).
struct NNPart { double d; int i; int pad; };

__global__ void __launch_bounds__(256) nn_exact_kernel(const NNArgs a, const int *list, const int *nlist, int *part_cnt, NNPart *parts)
{
    __shared__ double s_d[8];
    __shared__ int s_i[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = *nlist;
    const bool vec = (a.D & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.c0) | reinterpret_cast<uintptr_t>(a.c1)) & 15) == 0;
    // few undecided columns (the usual case: a fraction of a percent): a column's rows are split over up to 16 CTAs —
    // one CTA streaming all N rows of a column through L2 alone is latency-bound — and the CTA that finishes a column
    // last combines the parts
    int S = 1;
    if (n > 0 && n < (int)gridDim.x) { S = (int)gridDim.x / n; if (S > 16) S = 16; }
    const int rows_per = (a.N + S - 1) / S;
    for (int wi = blockIdx.x; wi < n * S; wi += gridDim.x) {
        const int u = wi / S, sp = wi - u * S;
        const long long w = list[u];
        const int pair = (int)(w / a.M), j = (int)(w % a.M);
        const float b = a.best_d[w];
        const float eps = (float)(a.D + 4) * 5.9604645e-8f;
        const float E = nn_margin_E(a, pair, w);
        const float bound = (b + E) * (1.0f + 4.0f * eps) + 1e-30f;
        const float *c0 = a.c0 + (size_t)pair * a.N * a.D;
        const float *q = a.c1 + ((size_t)pair * a.M + j) * a.D;
        double bd = __longlong_as_double(0x7ff0000000000000ll);
        int bi = 0x7fffffff;
        const int i_end = (sp + 1) * rows_per < a.N ? (sp + 1) * rows_per : a.N;
        for (int i = sp * rows_per + threadIdx.x; i < i_end; i += 256) {
            const float *p = c0 + (size_t)i * a.D;
            float acc32;
            if (vec) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 5
                for (int k = 0; k < a.D; k += 4) {
                    const float4 pv = __ldg(reinterpret_cast<const float4 *>(p + k));
                    const float4 qv = __ldg(reinterpret_cast<const float4 *>(q + k));
                    const float d0 = pv.x - qv.x, d1 = pv.y - qv.y, d2 = pv.z - qv.z, d3 = pv.w - qv.w;
                    s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
                }
                acc32 = (s0 + s1) + (s2 + s3);
            } else {
                acc32 = 0.0f;
                for (int k = 0; k < a.D; ++k) {
                    const float d = p[k] - q[k];
                    acc32 = fmaf(d, d, acc32);
                }
            }
            if (acc32 > bound) continue;
            double acc = 0.0;
#pragma unroll 4
            for (int k = 0; k < a.D; ++k) {
                double d = __dsub_rn((double)p[k], (double)q[k]);
                acc = __dadd_rn(acc, __dmul_rn(d, d));
            }
            double dist = __dsqrt_rn(acc);
            if (dist < bd) { bd = dist; bi = i; }  // ascending i within a thread: first minimum kept
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double od = __shfl_xor_sync(0xffffffffu, bd, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { s_d[warp] = bd; s_i[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < 8; ++k)
                if (s_d[k] < bd || (s_d[k] == bd && s_i[k] < bi)) { bd = s_d[k]; bi = s_i[k]; }
            if (S == 1) {
                a.out[w] = bi;     // ties -> lowest row, as numpy's argmin
            } else {
                NNPart *pp = parts + (size_t)u * 16;
                pp[sp].d = bd; pp[sp].i = bi;
                __threadfence();
                if (atomicAdd(part_cnt + u, 1) == S - 1) {           // the last part of this column
                    __threadfence();
                    for (int k = 0; k < S; ++k) {
                        const double od = *reinterpret_cast<volatile double *>(&pp[k].d);
                        const int oi = *reinterpret_cast<volatile int *>(&pp[k].i);
                        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                    }
                    a.out[w] = bi;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

static int nt_smem(int Kp) { return 6 * Kp * NT_B * 2 + 2 * NT_B * 4 + 64; }

int caelo_match_init(caelo_ctx *ctx)
{
    CAELO_CUDA(ctx, cudaFuncSetAttribute(nn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         2 * 256 * TILE * 4));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(nn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nt_smem(128)));
    return CAELO_OK;
}

extern "C" int caelo_nn_match(caelo_ctx *ctx, const float *codes0, const float *codes1, int P, int N,
                              int M, int D, int64_t *pair_idx, void *stream)
{
    if (!ctx || !codes0 || !codes1 || !pair_idx || P <= 0 || N <= 0 || M <= 0 || D <= 0 || D > 256)
        return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    size_t cols = (size_t)P * M, rows0 = (size_t)P * N;
    const int exact_grid = 4 * ctx->num_sms;
    int rc = caelo_reserve(ctx, ctx->misc, cols * 20 + rows0 * 4 + (size_t)P * 4 + 1024 + (size_t)exact_grid * (4 + 16 * sizeof(NNPart)));
    if (rc) return rc;
    NNArgs a;
    a.c0 = codes0; a.c1 = codes1; a.N = N; a.M = M; a.D = D; a.Dp = (D + 3) & ~3;
    a.best_d = reinterpret_cast<float *>(ctx->misc.ptr);
    a.second_d = a.best_d + cols;
    a.best_i = reinterpret_cast<int *>(a.second_d + cols);
    a.out = reinterpret_cast<long long *>(pair_idx);
    a.n1 = nullptr; a.nmax0 = nullptr; a.margin = ctx->nn_margin;
    const char *force = getenv("CAELO_NN_F32");   // debug switch: float32 CUDA-core pass for every D
    if (D <= 128 && !(force && force[0] == '1')) {
        const int Kp = (D + 15) & ~15, tiles0 = (N + NT_B - 1) / NT_B, tiles1 = (M + NT_B - 1) / NT_B;
        const size_t blk = (size_t)Kp * NT_B * 4;                     // one tile: hi + lo
        rc = caelo_reserve(ctx, ctx->match_ops, (size_t)P * (tiles0 + tiles1) * blk + (size_t)P * tiles0 * NT_B * 4 + 256);
        if (rc) return rc;
        unsigned char *ops0 = reinterpret_cast<unsigned char *>(ctx->match_ops.ptr), *ops1 = ops0 + (size_t)P * tiles0 * blk;
        float *n0p = reinterpret_cast<float *>(ops1 + (size_t)P * tiles1 * blk);
        float *n1 = reinterpret_cast<float *>(a.best_i + cols), *nmax = n1 + cols + rows0;
        CAELO_CUDA(ctx, caelo_fill_async(nmax, 0, (size_t)P * 4, st));
        { ProfScope ps_(ctx, "desc_prep_kernel", st);
          desc_prep_kernel<<<dim3(tiles0, P), NT_B, 0, st>>>(codes0, N, D, Kp, tiles0, ops0, nullptr, n0p, nmax);
          desc_prep_kernel<<<dim3(tiles1, P), NT_B, 0, st>>>(codes1, M, D, Kp, tiles1, ops1, n1, nullptr, nullptr); }
        CAELO_LAUNCH_CHECK(ctx);
        ctx->launches++;
        NNTcArgs t;
        t.ops0 = ops0; t.ops1 = ops1; t.n0p = n0p; t.n1 = n1; t.N = N; t.M = M; t.Kp = Kp; t.tiles0 = tiles0; t.tiles1 = tiles1;
        t.best_d = a.best_d; t.second_d = a.second_d; t.best_i = a.best_i;
        { ProfScope ps_(ctx, "nn_tc_kernel", st);
          nn_tc_kernel<<<dim3(tiles1, P), NT_THREADS, nt_smem(t.Kp), st>>>(t); }
        CAELO_LAUNCH_CHECK(ctx);
        a.n1 = n1; a.nmax0 = nmax;
    } else {
        dim3 grid((M + TILE - 1) / TILE, P);
        size_t smem = (size_t)2 * a.Dp * TILE * 4;
        { ProfScope ps_(ctx, "nn_tile_kernel", st); nn_tile_kernel<<<grid, MT_THREADS, smem, st>>>(a); }
        CAELO_LAUNCH_CHECK(ctx);
    }
    // undecided-column list lives behind every other array of the scratch block
    int *nlist = reinterpret_cast<int *>(reinterpret_cast<char *>(ctx->misc.ptr) + ((cols * 16 + rows0 * 4 + (size_t)P * 4 + 255) / 256) * 256);
    int *list = nlist + 16;
    int *part_cnt = list + ((cols + 15) / 16) * 16;
    NNPart *parts = reinterpret_cast<NNPart *>(part_cnt + ((exact_grid + 15) / 16) * 16);
    CAELO_CUDA(ctx, caelo_fill_async(nlist, 0, 64, st));
    CAELO_CUDA(ctx, caelo_fill_async(part_cnt, 0, (size_t)exact_grid * 4, st));
    { ProfScope ps_(ctx, "nn_decide_kernel", st); nn_decide_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(a, P, list, nlist); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "nn_exact_kernel", st); nn_exact_kernel<<<exact_grid, 256, 0, st>>>(a, list, nlist, part_cnt, parts); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}

// Test / measurement hook (tools/nn_margin.py): the approximate per-column results the LAST caelo_nn_match call left in the
// scratch (same P, N, M): best and second-best d^2 of the first pass, its best row, and how many columns went to the exact
// re-scan.  Any output pointer may be NULL.
extern "C" int caelo_debug_nn_last(caelo_ctx *ctx, int P, int N, int M, float *best_d, float *second_d, int32_t *best_i,
                                   int32_t *n_undecided, void *stream)
{
    if (!ctx || !ctx->misc.ptr || P <= 0 || N <= 0 || M <= 0) return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cols = (size_t)P * M, rows0 = (size_t)P * N;
    const float *bd = reinterpret_cast<const float *>(ctx->misc.ptr);
    const char *nl = reinterpret_cast<const char *>(ctx->misc.ptr) + ((cols * 16 + rows0 * 4 + (size_t)P * 4 + 255) / 256) * 256;
    if (best_d) CAELO_CUDA(ctx, cudaMemcpyAsync(best_d, bd, cols * 4, cudaMemcpyDeviceToDevice, st));
    if (second_d) CAELO_CUDA(ctx, cudaMemcpyAsync(second_d, bd + cols, cols * 4, cudaMemcpyDeviceToDevice, st));
    if (best_i) CAELO_CUDA(ctx, cudaMemcpyAsync(best_i, bd + 2 * cols, cols * 4, cudaMemcpyDeviceToDevice, st));
    if (n_undecided) CAELO_CUDA(ctx, cudaMemcpyAsync(n_undecided, nl, 4, cudaMemcpyDeviceToDevice, st));
    return CAELO_OK;
}

extern "C" int caelo_debug_set_nn_margin(caelo_ctx *ctx, float margin)
{
    if (!ctx || !(margin > 0.0f)) return CAELO_ERR_ARG;
    ctx->nn_margin = margin;
    return CAELO_OK;
}
