// extend.cu — ExtendKeyPtsInShpericalRing (reference SphericalRing.py:294-317) for sm_100a; the interest-point
// extension that feeds the pose refinement.
//
// The reference walks the key pixels in order; each takes every still-occupied pixel of its 13x13 window
// (row-major inside the window) and then ZEROES the counter window in place, so a pixel belongs to the FIRST
// key pixel whose window covers it.  Here: extend_claim_kernel takes atomicMin(key-pixel index) per covered
// pixel; extend_write_kernel (one CTA per frame, one warp per key pixel) counts each key pixel's owned,
// occupied pixels with ballots, scans the counts across the CTA and writes the points in the reference's
// order; extend_zero_kernel optionally applies the reference's in-place side effect on the counter.
#include "common.cuh"

namespace {

constexpr int RADIUS = 6;                         // nNeighborRadius (SphericalRing.py:295)
constexpr int NOOWNER = 0x7F7F7F7F;               // what cudaMemset(0x7F) leaves

struct ExtendArgs {
    const float *ring;        // [B,ring_H,ring_W,ring_C]
    void *counter;            // [B,cnt_H,cnt_W] int8 | int32
    const long long *kpix;    // [B,max_kpts,2] (row, col)
    const int *n_kpts;        // [B] or null (= max_kpts)
    int *owner;               // [B,cnt_H,cnt_W]
    float *ext;               // [B,ext_cap,3]
    int *n_ext;               // [B]
    int ring_C, ring_H, ring_W, cnt_kind, cnt_H, cnt_W, B, max_kpts, ext_cap;
};

// numpy slice [i-6 : i+7] of an axis of length n: a negative start wraps to n+start, which lies beyond the
// stop for every image this code sees -> empty window; the stop is clipped to n.
__device__ __forceinline__ bool window(const ExtendArgs &a, int row, int col, int &r0, int &r1, int &c0, int &c1)
{
    r0 = row - RADIUS; c0 = col - RADIUS;
    if (r0 < 0 || c0 < 0) return false;
    const int H = min(a.cnt_H, a.ring_H), W = min(a.cnt_W, a.ring_W);
    r1 = min(row + RADIUS + 1, H);
    c1 = min(col + RADIUS + 1, W);
    return r1 > r0 && c1 > c0;
}

__global__ void __launch_bounds__(256) extend_claim_kernel(const ExtendArgs a)
{
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int nk = a.n_kpts ? min(a.n_kpts[b], a.max_kpts) : a.max_kpts;
    if (k >= nk) return;
    const long long *px = a.kpix + ((size_t)b * a.max_kpts + k) * 2;
    int r0, r1, c0, c1;
    if (!window(a, (int)px[0], (int)px[1], r0, r1, c0, c1)) return;
    int *own = a.owner + (size_t)b * a.cnt_H * a.cnt_W;
    const int w = c1 - c0, n = (r1 - r0) * w;
    for (int j = lane; j < n; j += 32) atomicMin(own + (r0 + j / w) * a.cnt_W + c0 + j % w, k);
}

__device__ __forceinline__ bool occupied(const ExtendArgs &a, int b, int r, int c)
{
    const size_t i = ((size_t)b * a.cnt_H + r) * a.cnt_W + c;
    return a.cnt_kind == CAELO_COUNTER_I8 ? (reinterpret_cast<const int8_t *>(a.counter)[i] > 0)
                                          : (reinterpret_cast<const int32_t *>(a.counter)[i] > 0);
}

constexpr int EW_THREADS = 1024;
__global__ void __launch_bounds__(EW_THREADS) extend_write_kernel(const ExtendArgs a)
{
    extern __shared__ int s_cnt[];   // [max_kpts + 1] counts, then exclusive offsets
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nk = a.n_kpts ? min(a.n_kpts[b], a.max_kpts) : a.max_kpts;
    const int *own = a.owner + (size_t)b * a.cnt_H * a.cnt_W;
    const float *ring = a.ring + (size_t)b * a.ring_H * a.ring_W * a.ring_C;
    float *ext = a.ext + (size_t)b * a.ext_cap * 3;
    for (int pass = 0; pass < 2; ++pass) {
        for (int k = warp; k < nk; k += EW_THREADS / 32) {
            const long long *px = a.kpix + ((size_t)b * a.max_kpts + k) * 2;
            int r0, r1, c0, c1, total = 0;
            if (window(a, (int)px[0], (int)px[1], r0, r1, c0, c1)) {
                const int w = c1 - c0, n = (r1 - r0) * w;
                const int base = pass ? s_cnt[k] : 0;
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    bool take = false;
                    int r = 0, c = 0;
                    if (j < n) {
                        r = r0 + j / w; c = c0 + j % w;
                        take = own[r * a.cnt_W + c] == k && occupied(a, b, r, c);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, take);
                    if (pass && take) {
                        const int o = base + total + __popc(m & ((1u << lane) - 1));
                        if (o < a.ext_cap) {
                            const float *p = ring + ((size_t)r * a.ring_W + c) * a.ring_C;
                            ext[o * 3 + 0] = p[0]; ext[o * 3 + 1] = p[1]; ext[o * 3 + 2] = p[2];
                        }
                    }
                    total += __popc(m);
                }
            }
            if (!pass && lane == 0) s_cnt[k] = total;
        }
        __syncthreads();
        if (pass) break;
        // exclusive scan of s_cnt[0..nk) in place
        if (tid == 0) s_base = 0;
        __syncthreads();
        for (int t0 = 0; t0 < nk; t0 += EW_THREADS) {
            const int i = t0 + tid;
            const int c = i < nk ? s_cnt[i] : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                int w = s_warp[lane], wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += v;
                }
                s_warp[lane] = wi - w;
            }
            __syncthreads();
            const int excl = s_base + s_warp[warp] + inc - c;
            if (i < nk) s_cnt[i] = excl;
            __syncthreads();
            if (tid == EW_THREADS - 1) s_base = excl + c;
            __syncthreads();
        }
        if (tid == 0) a.n_ext[b] = s_base;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) extend_zero_kernel(const ExtendArgs a)
{
    const size_t n = (size_t)a.B * a.cnt_H * a.cnt_W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (a.owner[i] == NOOWNER) continue;                         // oneMask[:] = 0 (SphericalRing.py:307)
        if (a.cnt_kind == CAELO_COUNTER_I8) reinterpret_cast<int8_t *>(a.counter)[i] = 0;
        else reinterpret_cast<int32_t *>(a.counter)[i] = 0;
    }
}

}  // namespace

extern "C" int caelo_extend_keypoints(caelo_ctx *ctx, const float *ring, int ring_C, int ring_H, int ring_W, void *counter,
                                      int counter_dtype, int cnt_H, int cnt_W, const int64_t *kpix, const int32_t *n_kpts,
                                      int B, int max_kpts, float *ext, int ext_cap, int32_t *n_ext, int zero_counter,
                                      void *stream)
{
    if (!ctx || !ring || !counter || !kpix || !ext || !n_ext || B <= 0 || max_kpts <= 0 || ext_cap <= 0 || ring_C < 3)
        return CAELO_ERR_ARG;
    if (counter_dtype != CAELO_COUNTER_I8 && counter_dtype != CAELO_COUNTER_I32) return CAELO_ERR_ARG;
    if ((size_t)(max_kpts + 1) * 4 > 200 * 1024) return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npx = (size_t)B * cnt_H * cnt_W;
    int rc = caelo_reserve(ctx, ctx->scan_ws, npx * 4);
    if (rc) return rc;
    ExtendArgs a;
    a.ring = ring; a.counter = counter; a.kpix = reinterpret_cast<const long long *>(kpix); a.n_kpts = n_kpts;
    a.owner = reinterpret_cast<int *>(ctx->scan_ws.ptr);
    a.ext = ext; a.n_ext = n_ext;
    a.ring_C = ring_C; a.ring_H = ring_H; a.ring_W = ring_W; a.cnt_kind = counter_dtype; a.cnt_H = cnt_H; a.cnt_W = cnt_W;
    a.B = B; a.max_kpts = max_kpts; a.ext_cap = ext_cap;
    CAELO_CUDA(ctx, caelo_fill_async(a.owner, 0x7F, npx * 4, st));
    { ProfScope ps_(ctx, "extend_claim_kernel", st); extend_claim_kernel<<<dim3((max_kpts + 7) / 8, B), 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    const size_t smem = (size_t)(max_kpts + 1) * 4;
    if (smem > 48 * 1024)
        CAELO_CUDA(ctx, cudaFuncSetAttribute(extend_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { ProfScope ps_(ctx, "extend_write_kernel", st); extend_write_kernel<<<B, EW_THREADS, smem, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    if (zero_counter) {
        { ProfScope ps_(ctx, "extend_zero_kernel", st); extend_zero_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(a); }
        CAELO_LAUNCH_CHECK(ctx);
    }
    return CAELO_OK;
}
