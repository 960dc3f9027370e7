// respond_select.cu — a1 (RespondLayer.predict) and a2 (GetKeyPtsByAE) for sm_100a.
//
// Reference: /root/reference/SphericalRing.py:113-291 (selection), :405-408 (predict call),
// AE4SphericalRingPC.py:132-133 (the two conv layers).  Arithmetic contracts R1 and S1 are the
// ones written out in oracle/caelo_oracle.c; every float op that decides a keypoint index is an
// explicit round-to-nearest intrinsic so that no compiler flag can change the result.
//
// Kernels
//   respond_kernel        a1 alone: ring NHWC -> 8-channel response (the predict() boundary)
//   respond_score_kernel  <fused>: ring tile -> conv3x3+relu -> conv1x1+relu in shared memory
//                         -> 5x5 window-min score -> candidate keys; the response never
//                         reaches HBM.  <!fused>: same scoring from a response tile in HBM.
//   topk_kernel           per frame: radix-select the (maxk+1) largest (score,index) keys,
//                         bitonic-sort them, drop the best (quirk 1), emit points + pixels.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int TH = 16, TW = 64, HALO = 2;
constexpr int RH = TH + 2 * HALO, RW = TW + 2 * HALO;  // response region held per tile
constexpr int IH = RH + 2, IW = RW + 2;                // conv input region
constexpr int NPIX = RH * RW;
constexpr int kThreads = 256;
constexpr int EDGE = 8;  // Size4FilterTopEdge, SphericalRing.py:42

// contract R1 for one pixel; x[27] in (ky,kx,ci) order, out[8].  Scalar FFMA with the weight as a constant-bank
// operand: the packed FFMA2 form was measured 2x SLOWER here (0.40 -> 0.81 ms) — a 64-bit weight pair cannot be
// an immediate constant operand and has to come through uniform-register loads.  The channel loop is FULLY unrolled
// (1120 straight-line FFMAs, 18 KB of code): with `unroll 4` the weight offsets were run-time values and every FFMA
// was preceded by a uniform constant load (ncu source page, round 2: 93 M LDCU next to 119 M FFMA per launch).
// where the weights come from: the kernel-parameter constant bank (uniform loads) or a copy in shared memory
// (16-byte broadcast loads into ordinary registers)
struct WConst {
    const RespondWeights &w;
    __device__ __forceinline__ float4 w1(int t, int g) const { return make_float4(w.w1[t * 32 + 4 * g], w.w1[t * 32 + 4 * g + 1], w.w1[t * 32 + 4 * g + 2], w.w1[t * 32 + 4 * g + 3]); }
    __device__ __forceinline__ float4 b1(int g) const { return make_float4(w.b1[4 * g], w.b1[4 * g + 1], w.b1[4 * g + 2], w.b1[4 * g + 3]); }
    __device__ __forceinline__ float4 w2(int co, int h) const { return make_float4(w.w2[co * 8 + 4 * h], w.w2[co * 8 + 4 * h + 1], w.w2[co * 8 + 4 * h + 2], w.w2[co * 8 + 4 * h + 3]); }
    __device__ __forceinline__ float4 b2(int h) const { return make_float4(w.b2[4 * h], w.b2[4 * h + 1], w.b2[4 * h + 2], w.b2[4 * h + 3]); }
};
constexpr int WS_B1 = 27 * 32, WS_W2 = WS_B1 + 32, WS_B2 = WS_W2 + 256, WS_FLOATS = WS_B2 + 8;   // smem copy: w1 | b1 | w2 | b2
struct WSmem {
    const float *p;
    __device__ __forceinline__ float4 w1(int t, int g) const { return *reinterpret_cast<const float4 *>(p + t * 32 + 4 * g); }
    __device__ __forceinline__ float4 b1(int g) const { return *reinterpret_cast<const float4 *>(p + WS_B1 + 4 * g); }
    __device__ __forceinline__ float4 w2(int co, int h) const { return *reinterpret_cast<const float4 *>(p + WS_W2 + co * 8 + 4 * h); }
    __device__ __forceinline__ float4 b2(int h) const { return *reinterpret_cast<const float4 *>(p + WS_B2 + 4 * h); }
};

template <int PPT, class W>
__device__ __forceinline__ void respond_pixels(const W &w, const float (&x)[PPT][27], float (&out)[PPT][8])
{
    // PPT pixels per thread share every weight load, and the 32 hidden channels are walked four at a time with the tap
    // loop inside, so that the four weights of a tap (contiguous in w1[t][co]) come in with one 16-byte load.
    // Every per-channel chain is still bias, then fmaf over the taps in ascending order, and the 1x1 layer still
    // accumulates the hidden channels in ascending order (contract R1).
    {
        const float4 c0 = w.b2(0), c1 = w.b2(1);
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            out[p][0] = c0.x; out[p][1] = c0.y; out[p][2] = c0.z; out[p][3] = c0.w;
            out[p][4] = c1.x; out[p][5] = c1.y; out[p][6] = c1.z; out[p][7] = c1.w;
        }
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        float acc[PPT][4];
        const float4 bb = w.b1(g);
#pragma unroll
        for (int p = 0; p < PPT; ++p) { acc[p][0] = bb.x; acc[p][1] = bb.y; acc[p][2] = bb.z; acc[p][3] = bb.w; }
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            const float4 wv = w.w1(t, g);
            const float wk[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int p = 0; p < PPT; ++p) acc[p][k] = __fmaf_rn(x[p][t], wk[k], acc[p][k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 u0 = w.w2(4 * g + k, 0), u1 = w.w2(4 * g + k, 1);
            const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
            for (int c2 = 0; c2 < 8; ++c2)
#pragma unroll
                for (int p = 0; p < PPT; ++p) out[p][c2] = __fmaf_rn(fmaxf(acc[p][k], 0.0f), uu[c2], out[p][c2]);
        }
    }
#pragma unroll
    for (int p = 0; p < PPT; ++p)
#pragma unroll
        for (int c2 = 0; c2 < 8; ++c2) out[p][c2] = fmaxf(out[p][c2], 0.0f);
}

template <int P, int kVar>
__device__ __forceinline__ void conv_px(const RespondWeights &w, const float *w_s, const float (&x)[P][27], float (&out)[P][8])
{
    if constexpr (kVar >= 2) respond_pixels<P>(WSmem{w_s}, x, out);
    else respond_pixels<P>(WConst{w}, x, out);
}

__device__ __forceinline__ void respond_pixel(const RespondWeights &w, const float (&x)[27], float (&out)[8])
{
    respond_pixels<1>(WConst{w}, reinterpret_cast<const float (&)[1][27]>(x), reinterpret_cast<float (&)[1][8]>(out));
}

__global__ void __launch_bounds__(kThreads)
respond_kernel(const __grid_constant__ RespondWeights w, const float *__restrict__ ring, int B,
               int H, int W, float *__restrict__ resp)
{
    const long long total = (long long)B * H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        int c = (int)(p % W);
        int r = (int)((p / W) % H);
        const float *img = ring + (p / ((long long)H * W)) * (long long)H * W * 3;
        float x[27];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                int rr = r + ky - 1, cc = c + kx - 1;
                bool in = rr >= 0 && rr < H && cc >= 0 && cc < W;
                const float *q = img + ((long long)rr * W + cc) * 3;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) x[(ky * 3 + kx) * 3 + ci] = in ? __ldg(q + ci) : 0.0f;
            }
        float out[8];
        respond_pixel(w, x, out);
        float4 *o = reinterpret_cast<float4 *>(resp + p * 8);
        o[0] = make_float4(out[0], out[1], out[2], out[3]);
        o[1] = make_float4(out[4], out[5], out[6], out[7]);
    }
}

struct SelectArgs {
    const float *ring;   // [B,ring_H,ring_W,ring_C]
    const void *counter; // [B,cnt_H,cnt_W]
    const float *resp;   // [B,H,W,8] (unfused mode)
    unsigned long long *keys;  // [B,H*W]
    int *count;                // [B]
    int ring_C, ring_H, ring_W, cnt_kind, cnt_H, cnt_W, H, W, B;
};

__device__ __forceinline__ bool occupied(const SelectArgs &a, int b, int r, int c)
{
    if (r < 0 || r >= a.cnt_H || c < 0 || c >= a.cnt_W) return false;
    size_t i = ((size_t)b * a.cnt_H + r) * a.cnt_W + c;
    return a.cnt_kind == CAELO_COUNTER_I8 ? (reinterpret_cast<const int8_t *>(a.counter)[i] > 0)
                                          : (reinterpret_cast<const int32_t *>(a.counter)[i] > 0);
}

// kVar (fused only): 0 = one pixel per thread and round, weights from the constant bank; 1 = two pixels, constant bank;
// 2 = two pixels, weights from shared memory; 3 = one pixel, shared memory
template <bool kFused, int kVar>
__global__ void __launch_bounds__(kThreads, !kFused ? 1 : ((kVar == 0 || kVar == 3) ? 3 : 2))
respond_score_kernel(const __grid_constant__ RespondWeights w, const SelectArgs a)
{
    extern __shared__ float smem[];
    __shared__ __align__(16) float w_s[(kFused && kVar >= 2) ? WS_FLOATS : 4];
    if (kFused && kVar >= 2) {
        for (int i = threadIdx.x; i < WS_FLOATS; i += kThreads)
            w_s[i] = i < WS_B1 ? w.w1[i] : (i < WS_W2 ? w.b1[i - WS_B1] : (i < WS_B2 ? w.w2[i - WS_W2] : w.b2[i - WS_B2]));
    }
    float *resp_s = smem;                                   // [8][NPIX]
    unsigned char *occ_s = reinterpret_cast<unsigned char *>(resp_s + 8 * NPIX);  // [NPIX]
    float *in_s = reinterpret_cast<float *>(occ_s + ((NPIX + 15) / 16) * 16);      // [IH*IW*3]

    const int b = blockIdx.z;
    const int r0 = EDGE + blockIdx.y * TH - HALO;  // image row of region row 0
    const int c0 = EDGE + blockIdx.x * TW - HALO;
    const int H = a.H, W = a.W;
    const float *ring_b = a.ring + (size_t)b * a.ring_H * a.ring_W * a.ring_C;

    if (kFused) {
        // stage the conv input (zero padded outside the H x W image)
        for (int i = threadIdx.x; i < IH * IW; i += kThreads) {
            int rr = r0 - 1 + i / IW, cc = c0 - 1 + i % IW;
            bool in = rr >= 0 && rr < H && cc >= 0 && cc < W;
            const float *q = ring_b + ((size_t)rr * a.ring_W + cc) * a.ring_C;
            in_s[i * 3 + 0] = in ? __ldg(q + 0) : 0.0f;
            in_s[i * 3 + 1] = in ? __ldg(q + 1) : 0.0f;
            in_s[i * 3 + 2] = in ? __ldg(q + 2) : 0.0f;
        }
        __syncthreads();
    }
    // ---- occupancy of the region + two compact work lists: occupied pixels (the only ones whose response is
    //      ever read: as a centre or as a valid neighbour) and interior pixels that can be key points ----
    unsigned short *list_occ = reinterpret_cast<unsigned short *>(in_s + (kFused ? IH * IW * 3 : 0));  // [NPIX]
    unsigned short *list_ctr = list_occ + NPIX;                                                        // [TH*TW]
    __shared__ int s_nocc, s_nctr;
    if (threadIdx.x == 0) { s_nocc = 0; s_nctr = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int i0 = 0; i0 < NPIX; i0 += kThreads) {   // uniform trip count: the ballots need every lane
        const int i = i0 + threadIdx.x;
        bool occ = false, ctr = false;
        if (i < NPIX) {
            const int lr = i / RW, lc = i % RW;
            const int rr = r0 + lr, cc = c0 + lc;
            const bool in = rr >= 0 && rr < H && cc >= 0 && cc < W;
            occ = in && occupied(a, b, rr, cc);
            occ_s[i] = occ ? 1 : 0;
            // rows [8,H-8); cols [8,W-8) from the final filter; quirk 2 removes cols [H-8,H)
            ctr = occ && lr >= HALO && lr < HALO + TH && lc >= HALO && lc < HALO + TW && rr >= EDGE && rr < H - EDGE &&
                  cc >= EDGE && cc < W - EDGE && !(cc >= H - EDGE && cc < H);
        }
        const unsigned mo = __ballot_sync(0xffffffffu, occ), mc = __ballot_sync(0xffffffffu, ctr);
        int bo = 0, bc = 0;
        if (lane == 0) {
            if (mo) bo = atomicAdd(&s_nocc, __popc(mo));
            if (mc) bc = atomicAdd(&s_nctr, __popc(mc));
        }
        bo = __shfl_sync(0xffffffffu, bo, 0);
        bc = __shfl_sync(0xffffffffu, bc, 0);
        if (occ) list_occ[bo + __popc(mo & ((1u << lane) - 1u))] = (unsigned short)i;
        if (ctr) list_ctr[bc + __popc(mc & ((1u << lane) - 1u))] = (unsigned short)i;
    }
    __syncthreads();
    const int nocc = s_nocc, nctr = s_nctr;
    // ---- response of the occupied pixels (fused: two listed pixels per thread and round, see respond_pixels) ----
    if (kFused) {
        constexpr int PPT = (kVar == 1 || kVar == 2) ? 2 : 1;
        for (int n0 = 0; n0 < nocc; n0 += PPT * kThreads) {
            const int na = n0 + threadIdx.x, nb = na + kThreads;
            if (na >= nocc) break;
            const int ia = list_occ[na], ib = (PPT == 2) ? list_occ[nb < nocc ? nb : na] : ia;
            float x[PPT][27], out[PPT][8];
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const int i = p ? ib : ia, lr = i / RW, lc = i % RW;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int ci = 0; ci < 3; ++ci)
                            x[p][(ky * 3 + kx) * 3 + ci] = in_s[((lr + ky) * IW + (lc + kx)) * 3 + ci];
            }
            if (PPT == 2 && nb < nocc) {
                conv_px<PPT, kVar>(w, w_s, x, out);
#pragma unroll
                for (int k = 0; k < 8; ++k) { resp_s[k * NPIX + ia] = out[0][k]; resp_s[k * NPIX + ib] = out[PPT - 1][k]; }
            } else {
                conv_px<1, kVar>(w, w_s, reinterpret_cast<const float (&)[1][27]>(x), reinterpret_cast<float (&)[1][8]>(out));
#pragma unroll
                for (int k = 0; k < 8; ++k) resp_s[k * NPIX + ia] = out[0][k];
            }
        }
    } else {
        for (int n = threadIdx.x; n < nocc; n += kThreads) {
            const int i = list_occ[n];
            const int lr = i / RW, lc = i % RW;
            const float4 *q = reinterpret_cast<const float4 *>(
                a.resp + (((size_t)b * H + (r0 + lr)) * W + (c0 + lc)) * 8);
            float4 v0 = __ldg(q), v1 = __ldg(q + 1);
            resp_s[0 * NPIX + i] = v0.x; resp_s[1 * NPIX + i] = v0.y; resp_s[2 * NPIX + i] = v0.z; resp_s[3 * NPIX + i] = v0.w;
            resp_s[4 * NPIX + i] = v1.x; resp_s[5 * NPIX + i] = v1.y; resp_s[6 * NPIX + i] = v1.z; resp_s[7 * NPIX + i] = v1.w;
        }
    }
    __syncthreads();

    // ---- score the listed centres (contract S1) ----
    for (int n0 = 0; n0 < nctr; n0 += kThreads) {   // uniform trip count for the ballot below
        const int n = n0 + threadIdx.x;
        bool keep = false;
        float best = __int_as_float(0x7f800000);
        int r = 0, c = 0;
        if (n < nctr) {
            const int li = list_ctr[n];
            const int lr = li / RW, lc = li % RW;
            r = r0 + lr; c = c0 + lc;
            float ctr[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) ctr[k] = resp_s[k * NPIX + li];
            int count = 0;
#pragma unroll
            for (int dr = -2; dr <= 2; ++dr)
#pragma unroll
                for (int dc = -2; dc <= 2; ++dc) {
                    if (dr == 0 && dc == 0) continue;
                    int ni = li + dr * RW + dc;
                    if (!occ_s[ni]) continue;
                    float sq[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float d = __fsub_rn(resp_s[k * NPIX + ni], ctr[k]);
                        sq[k] = __fmul_rn(d, d);
                    }
                    float s = __fadd_rn(__fadd_rn(__fadd_rn(sq[0], sq[1]), __fadd_rn(sq[2], sq[3])),
                                        __fadd_rn(__fadd_rn(sq[4], sq[5]), __fadd_rn(sq[6], sq[7])));
                    float nrm = __fsqrt_rn(s);
                    best = fminf(best, nrm);
                    ++count;
                }
            if (count >= 5 && (double)best > 0.2) {
                const float *p = ring_b + ((size_t)r * a.ring_W + c) * a.ring_C;
                float s = 0.0f;
                for (int k = 0; k < a.ring_C; ++k) {
                    float v = __ldg(p + k);
                    s = __fadd_rn(s, __fmul_rn(v, v));
                }
                keep = __fsqrt_rn(s) >= 10.0f;  // VisibleBottom, SphericalRing.py:39,197-198
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(a.count + b, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (keep) {
                unsigned long long key =
                    ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)(r * W + c);
                a.keys[(size_t)b * H * W + base + __popc(m & ((1u << lane) - 1u))] = key;
            }
        }
    }
}

// ---- top-k ------------------------------------------------------------------------------
struct TopkArgs {
    const unsigned long long *keys;  // [B,cap]
    const int *count;                // [B]
    const float *ring;
    int ring_C, ring_H, ring_W, W, cap, maxk, sort_n;
    float *kpts;        // [B,maxk,3]
    long long *kpix;    // [B,maxk,2]
    int *n_kpts;        // [B]
};

constexpr int TOPK_CAND = 3072;   // keys tied with the 16-bit prefix that fit the shared-memory stage

__global__ void __launch_bounds__(1024) topk_kernel(const TopkArgs a)
{
    extern __shared__ unsigned long long sk[];  // [sort_n]
    __shared__ int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_krem, s_fill, s_surv, s_ncand;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = a.count[b];
    const unsigned long long *keys = a.keys + (size_t)b * a.cap;
    const int want = a.maxk + 1;  // keep maxk+1 largest, then drop the best (quirk 1)
    const int S = a.sort_n;

    for (int i = tid; i < S; i += blockDim.x) sk[i] = 0ull;
    if (tid == 0) s_fill = 0;
    __syncthreads();

    if (n <= want) {
        for (int i = tid; i < n; i += blockDim.x) sk[i] = keys[i];
    } else {
        // MSB-first radix select of the want-th largest key (keys are unique).  The first two digits (16 bits of
        // the score) are counted over all n keys in global memory; then the keys still tied with the prefix —
        // usually a few dozen — are compacted into shared memory (the ones above it go straight to the output),
        // and the remaining six digits and the final collection only touch those.
        unsigned long long *cand = sk + S;   // [TOPK_CAND]
        if (tid == 0) { s_prefix = 0ull; s_krem = want; s_ncand = 0; }
        const unsigned long long *src = keys;
        int nsrc = n;
        bool compact = false;
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            if (pass == 2 && s_surv <= TOPK_CAND) {
                const unsigned long long p16 = s_prefix >> 48;
                for (int i0 = 0; i0 < n; i0 += blockDim.x) {   // uniform trip count for the ballots
                    const int i = i0 + tid;
                    const unsigned long long k = i < n ? keys[i] : 0ull;
                    const bool above = i < n && (k >> 48) > p16, tied = i < n && (k >> 48) == p16;
                    const unsigned ma = __ballot_sync(0xffffffffu, above), mt = __ballot_sync(0xffffffffu, tied);
                    int ba = 0, bt = 0;
                    if ((tid & 31) == 0) {
                        if (ma) ba = atomicAdd(&s_fill, __popc(ma));
                        if (mt) bt = atomicAdd(&s_ncand, __popc(mt));
                    }
                    ba = __shfl_sync(0xffffffffu, ba, 0);
                    bt = __shfl_sync(0xffffffffu, bt, 0);
                    const unsigned below = (1u << (tid & 31)) - 1u;
                    if (above) { const int pos = ba + __popc(ma & below); if (pos < S) sk[pos] = k; }
                    if (tied) cand[bt + __popc(mt & below)] = k;
                }
                __syncthreads();
                src = cand; nsrc = s_ncand; compact = true;
            }
            const unsigned long long prefix = s_prefix;
            const unsigned long long hmask = pass == 0 ? 0ull : (~0ull << (shift + 8));
            // the leading digits of the scores are nearly constant (same exponent): aggregate equal digits inside
            // a warp so that one shared-memory atomic stands for up to 32 keys (uniform trip count for the match)
            for (int i0 = 0; i0 < nsrc; i0 += blockDim.x) {
                const int i = i0 + tid;
                int digit = -1;
                if (i < nsrc) {
                    unsigned long long k = src[i];
                    if ((k & hmask) == prefix) digit = (int)((k >> shift) & 0xff);
                }
                const unsigned grp = __match_any_sync(0xffffffffu, digit);
                if (digit >= 0 && (tid & 31) == __ffs(grp) - 1) atomicAdd(&hist[digit], __popc(grp));
            }
            __syncthreads();
            if (tid < 32) {
                // descending scan for the digit holding the krem-th largest key: lane L owns digits 255-8L .. 248-8L
                int c[8], tot = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * tid - j]; tot += c[j]; }
                int incl = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += v;
                }
                const int krem = s_krem;
                int cum = incl - tot;
                if (cum < krem && krem <= incl) {
                    int d = 255 - 8 * tid, surv = c[7];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (cum + c[j] >= krem) { surv = c[j]; break; }
                        cum += c[j];
                        --d;
                    }
                    s_krem = krem - cum;
                    s_prefix = prefix | ((unsigned long long)d << shift);
                    s_surv = surv;       // keys that share the digits chosen so far
                }
            }
            __syncthreads();
        }
        const unsigned long long kth = s_prefix;
        // collect every key >= kth (with the compacted source the keys above the 16-bit prefix are already in)
        for (int i0 = 0; i0 < nsrc; i0 += blockDim.x) {   // uniform trip count for the ballot
            const int i = i0 + tid;
            const unsigned long long k = i < nsrc ? src[i] : 0ull;
            const bool take = i < nsrc && k >= kth;
            const unsigned m = __ballot_sync(0xffffffffu, take);
            int base = 0;
            if (m && (tid & 31) == 0) base = atomicAdd(&s_fill, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (take) {
                const int pos = base + __popc(m & ((1u << (tid & 31)) - 1u));
                if (pos < S) sk[pos] = k;
            }
        }
        (void)compact;
    }
    __syncthreads();

    // bitonic sort ascending; zero padding sinks to the front
    for (int k = 2; k <= S; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < S; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long x = sk[i], y = sk[ixj];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { sk[i] = y; sk[ixj] = x; }
                }
            }
            __syncthreads();
        }

    const int have = n < want ? n : want;
    const int nout = have > 0 ? have - 1 : 0;   // candidates[-maxk-1:-1]
    const float *ring_b = a.ring + (size_t)b * a.ring_H * a.ring_W * a.ring_C;
    for (int i = tid; i < a.maxk; i += blockDim.x) {
        float x = 0.f, y = 0.f, z = 0.f;
        long long r = 0, c = 0;
        if (i < nout) {
            unsigned idx = (unsigned)(sk[S - 1 - nout + i] & 0xffffffffu);
            r = idx / (unsigned)a.W;
            c = idx % (unsigned)a.W;
            const float *p = ring_b + ((size_t)r * a.ring_W + c) * a.ring_C;
            x = p[0]; y = p[1]; z = p[2];
        }
        float *o = a.kpts + ((size_t)b * a.maxk + i) * 3;
        o[0] = x; o[1] = y; o[2] = z;
        long long *q = a.kpix + ((size_t)b * a.maxk + i) * 2;
        q[0] = r; q[1] = c;
    }
    if (tid == 0) a.n_kpts[b] = nout;
}

int launch_select(caelo_ctx *ctx, bool fused, const float *resp, int H, int W, const float *ring,
                  int ring_C, int ring_H, int ring_W, const void *counter, int counter_dtype,
                  int cnt_H, int cnt_W, int B, int max_kpts, float *kpts, int64_t *kpix,
                  int32_t *n_kpts, cudaStream_t st)
{
    if (!ring || !counter || !kpts || !kpix || !n_kpts || B <= 0) return CAELO_ERR_ARG;
    if (max_kpts <= 0 || max_kpts > 4095) return CAELO_ERR_ARG;
    if (H < 2 * EDGE + 1 || W < 2 * EDGE + 1 || ring_H < H || ring_W < W || ring_C < 3 || ring_C > 8)
        return CAELO_ERR_ARG;
    if (counter_dtype != CAELO_COUNTER_I8 && counter_dtype != CAELO_COUNTER_I32) return CAELO_ERR_ARG;
    if (fused && !ctx->have_respond) return CAELO_ERR_NO_WEIGHTS;
    const size_t cap = (size_t)H * W;
    size_t need = (size_t)B * cap * 8 + (size_t)B * 4 + 256;
    int rc = caelo_reserve(ctx, ctx->cand, need);
    if (rc) return rc;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ctx->cand.ptr);
    int *count = reinterpret_cast<int *>(keys + (size_t)B * cap);
    CAELO_CUDA(ctx, caelo_fill_async(count, 0, (size_t)B * 4, st));

    SelectArgs a;
    a.ring = ring; a.counter = counter; a.resp = resp; a.keys = keys; a.count = count;
    a.ring_C = ring_C; a.ring_H = ring_H; a.ring_W = ring_W; a.cnt_kind = counter_dtype;
    a.cnt_H = cnt_H; a.cnt_W = cnt_W; a.H = H; a.W = W; a.B = B;
    dim3 grid((W - 2 * EDGE + TW - 1) / TW, (H - 2 * EDGE + TH - 1) / TH, B);
    size_t smem = (size_t)8 * NPIX * 4 + ((NPIX + 15) / 16) * 16 + (fused ? (size_t)IH * IW * 3 * 4 : 0) + (NPIX + TH * TW) * 2;
    if (fused) {
        static int var = -1;
        if (var < 0) { const char *e = getenv("CAELO_RESPOND_VARIANT"); var = e ? atoi(e) : 0; }
        ProfScope ps_(ctx, "respond_score_kernel<fused>", st);
        switch (var) {
        case 1: respond_score_kernel<true, 1><<<grid, kThreads, smem, st>>>(ctx->respond_host, a); break;
        case 3: respond_score_kernel<true, 3><<<grid, kThreads, smem, st>>>(ctx->respond_host, a); break;
        case 2: respond_score_kernel<true, 2><<<grid, kThreads, smem, st>>>(ctx->respond_host, a); break;
        default: respond_score_kernel<true, 0><<<grid, kThreads, smem, st>>>(ctx->respond_host, a); break;
        }
    } else {
        if (!resp) return CAELO_ERR_ARG;
        { ProfScope ps_(ctx, "respond_score_kernel<from_resp>", st); respond_score_kernel<false, 0><<<grid, kThreads, smem, st>>>(ctx->respond_host, a); }
    }
    CAELO_LAUNCH_CHECK(ctx);

    TopkArgs t;
    t.keys = keys; t.count = count; t.ring = ring; t.ring_C = ring_C; t.ring_H = ring_H;
    t.ring_W = ring_W; t.W = W; t.cap = (int)cap; t.maxk = max_kpts;
    int S = 2;
    while (S < max_kpts + 1) S <<= 1;
    t.sort_n = S;
    t.kpts = kpts; t.kpix = reinterpret_cast<long long *>(kpix); t.n_kpts = n_kpts;
    { ProfScope ps_(ctx, "topk_kernel", st); topk_kernel<<<B, 1024, (size_t)(S + TOPK_CAND) * 8, st>>>(t); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}

}  // namespace

int caelo_select_init(caelo_ctx *ctx)
{
    const int smem = 8 * NPIX * 4 + ((NPIX + 15) / 16) * 16 + IH * IW * 3 * 4 + (NPIX + TH * TW) * 2;
    CAELO_CUDA(ctx, cudaFuncSetAttribute(respond_score_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(respond_score_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(respond_score_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(respond_score_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(respond_score_kernel<false, 0>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (4096 + TOPK_CAND) * 8));
    return CAELO_OK;
}

extern "C" int caelo_respond_forward(caelo_ctx *ctx, const float *ring, int B, int H, int W,
                                     float *resp, void *stream)
{
    if (!ctx || !ring || !resp || B <= 0 || H <= 0 || W <= 0) return CAELO_ERR_ARG;
    if (!ctx->have_respond) return CAELO_ERR_NO_WEIGHTS;
    long long total = (long long)B * H * W;
    long long blocks = (total + kThreads - 1) / kThreads;
    long long maxb = (long long)ctx->num_sms * 32;
    if (blocks > maxb) blocks = maxb;
    { ProfScope ps_(ctx, "respond_kernel", (cudaStream_t)stream); respond_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(ctx->respond_host, ring, B, H, W, resp); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}

extern "C" int caelo_select_keypoints(caelo_ctx *ctx, const float *resp, int H, int W,
                                      const float *ring, int ring_C, int ring_H, int ring_W,
                                      const void *counter, int counter_dtype, int cnt_H, int cnt_W,
                                      int B, int max_kpts, float *kpts, int64_t *kpix,
                                      int32_t *n_kpts, void *stream)
{
    if (!ctx) return CAELO_ERR_ARG;
    return launch_select(ctx, false, resp, H, W, ring, ring_C, ring_H, ring_W, counter, counter_dtype,
                         cnt_H, cnt_W, B, max_kpts, kpts, kpix, n_kpts, (cudaStream_t)stream);
}

extern "C" int caelo_respond_select(caelo_ctx *ctx, const float *ring, int ring_C, int ring_H,
                                    int ring_W, const void *counter, int counter_dtype, int cnt_H,
                                    int cnt_W, int H, int W, int B, int max_kpts, float *kpts,
                                    int64_t *kpix, int32_t *n_kpts, float *resp_out, void *stream)
{
    if (!ctx) return CAELO_ERR_ARG;
    if (resp_out) {
        // the caller wants the response image as well: a1 to HBM, then score from it
        if (ring_C != 3 || ring_H != H || ring_W != W) return CAELO_ERR_UNSUPPORTED;
        int rc = caelo_respond_forward(ctx, ring, B, H, W, resp_out, stream);
        if (rc) return rc;
        return launch_select(ctx, false, resp_out, H, W, ring, ring_C, ring_H, ring_W, counter,
                             counter_dtype, cnt_H, cnt_W, B, max_kpts, kpts, kpix, n_kpts,
                             (cudaStream_t)stream);
    }
    return launch_select(ctx, true, nullptr, H, W, ring, ring_C, ring_H, ring_W, counter, counter_dtype,
                         cnt_H, cnt_W, B, max_kpts, kpts, kpix, n_kpts, (cudaStream_t)stream);
}
