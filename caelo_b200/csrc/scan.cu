// scan.cu — the two per-scan pre-stages in front of the hot path for sm_100a.
//
//   f1  ProjectPC2SphericalRing (reference SphericalRing.py:72-94): scan (N,4) -> 69x1800x5 ring image +
//       hit counter.  The reference walks the points in file order, so the LAST point that lands in a
//       pixel owns it: ring_claim_kernel takes atomicMax(point index) per pixel (+ atomicAdd on the
//       counter), ring_fill_kernel then writes every pixel of every requested output exactly once
//       (coalesced, no memset of the outputs).  Arithmetic contract P1 (oracle/caelo_oracle.c):
//       float32 norm without contraction, float32 z/r, float64 atan2/asin/divide, C truncation.
//
//   f2  Voxelization (reference Voxel.py:89-173): scan -> three ordered occupied-voxel lists (2 cm, 16 cm,
//       64 cm) + block list / block offsets / in-block coordinates.  The reference's python loop defines
//       the ORDER: lists 1 and 2 in first-seen order, list 0 grouped by 1.28 m block in block-first-seen
//       order and first-seen inside a block.  Here every voxel / block key goes into an open-addressing
//       hash table that keeps the smallest point index (voxel_claim_kernel); a point "wins" a key iff it
//       is that smallest index, and an ordered stream compaction of the winners (voxel_compact_kernel,
//       one CTA per frame and table, ballot/shuffle scans) IS the first-seen order.  List 0 additionally
//       needs a stable grouping by block rank: per-block counts -> exclusive scan -> unordered slot claim
//       -> rank-by-counting inside each (short) block segment.  Integer-exact; float64 only for the
//       (p + Visible) / size divides, contract V1.
#include "common.cuh"
#include "voxel_math.cuh"

namespace {

constexpr unsigned long long EMPTY = ~0ull;
constexpr unsigned NOSLOT = 0xFFFFFFFFu;
constexpr int RING_H = 69, RING_W = 1800;       // ImgH, ImgW (SphericalRing.py:60-61)
constexpr int CNN_H = 64, CNN_W = 1792;         // nLines, ImgW - CropWidth_SphericalRing

__device__ __forceinline__ unsigned hash64(unsigned long long k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned)k;
}

// ---------------------------------------------------------------------------------------------------
// f1
// ---------------------------------------------------------------------------------------------------
struct RingArgs {
    const float *pts;            // rows of 4
    const long long *offsets;    // dev [F+1]
    int *winner;                 // [F,69,1800] last point index (local to the frame) or -1
    int *hits;                   // [F,69,1800]
    int *status;                 // [F] or null: points with column == ImgW
    double az_res, v_res, v_off;
    float *ring5;
    int *counter_i32;
    float *ring3;
    signed char *counter_i8;
    int F;
};

__device__ __forceinline__ float norm3_p1(float x, float y, float z)
{
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

__global__ void __launch_bounds__(256) ring_claim_kernel(const RingArgs a)
{
    const int f = blockIdx.y;
    const long long beg = a.offsets[f], n = a.offsets[f + 1] - beg;
    const float4 *p4 = reinterpret_cast<const float4 *>(a.pts) + beg;
    int *win = a.winner + (size_t)f * RING_H * RING_W, *hits = a.hits + (size_t)f * RING_H * RING_W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float4 p = p4[i];
        const float r = norm3_p1(p.x, p.y, p.z);
        if (!(r > 0.0f)) continue;                                   // SphericalRing.py:77-80
        const int col = (int)__ddiv_rn(__dsub_rn(3.14159265358979323846, atan2((double)p.y, (double)p.x)), a.az_res);
        const double beta = asin((double)__fdiv_rn(p.z, r));
        const int row = RING_H - (int)__dadd_rn(__ddiv_rn(beta, a.v_res), a.v_off);
        if (row < 0 || row >= RING_H) continue;                      // :89-90
        if (col < 0 || col >= RING_W) {                              // numpy raises IndexError here
            if (a.status) atomicAdd(a.status + f, 1);
            continue;
        }
        atomicMax(win + row * RING_W + col, (int)i);                 // last point in file order wins (:91-92)
        atomicAdd(hits + row * RING_W + col, 1);                     // :93
    }
}

__global__ void __launch_bounds__(256) ring_fill_kernel(const RingArgs a)
{
    const int f = blockIdx.y;
    const float4 *p4 = reinterpret_cast<const float4 *>(a.pts) + a.offsets[f];
    const size_t fo = (size_t)f * RING_H * RING_W;
    for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < RING_H * RING_W; px += gridDim.x * blockDim.x) {
        const int w = a.winner[fo + px], h = a.hits[fo + px];
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        float r = 0.f;
        if (w >= 0) {
            p = p4[w];
            r = norm3_p1(p.x, p.y, p.z);
        }
        if (a.ring5) {
            float *o = a.ring5 + (fo + px) * 5;
            o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.w; o[4] = r;
        }
        if (a.counter_i32) a.counter_i32[fo + px] = h;
        if (a.counter_i8) a.counter_i8[fo + px] = (signed char)h;   // numpy's wrapping cast (BatchPreprocess.py:107)
        if (a.ring3) {
            const int row = px / RING_W, col = px - row * RING_W;
            if (row < CNN_H && col < CNN_W) {
                float *o = a.ring3 + (((size_t)f * CNN_H + row) * CNN_W + col) * 3;
                o[0] = p.x; o[1] = p.y; o[2] = p.z;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// f2
// ---------------------------------------------------------------------------------------------------
struct VoxArgs {
    const float *pts;
    const long long *offsets;    // dev [F+1]
    unsigned long long *keys;    // [F,4,capT]
    int *vals;                   // [F,4,capT] smallest point index per key
    unsigned *brank;             // [F,capT] block-table slot -> block rank
    unsigned *slots;             // [4, total_points] slot of each point's key in table t (NOSLOT = filtered out)
    long long total_points;
    unsigned capT;               // power of two
    // compaction products, per frame stride = cap
    unsigned *w0_slot, *w0_bslot, *w0_blk, *seg, *seg_blk;
    int *bcount, *bfill;         // [F,cap]
    // outputs
    short *vox;                  // [F,3,cap,3]
    short *local0;               // [F,cap,3] or null
    short *blocks;               // [F,cap,3] or null
    int *cnt;                    // [F,cap+1] (scratch if the caller passes null)
    int *counts;                 // [F,4]
    int *status;                 // [F] or null
    int cap, F;
};

__device__ __forceinline__ unsigned long long vkey(int x, int y, int z)
{
    return (unsigned long long)(unsigned)x | ((unsigned long long)(unsigned)y << 16) | ((unsigned long long)(unsigned)z << 32);
}

// Key -> slot with the smallest point index kept in vals.  Plain loads first: most points of a scan repeat a
// key that is already present with a smaller index, so neither the CAS nor the atomicMin is issued for them.
__device__ __forceinline__ unsigned claim(unsigned long long *keys, int *vals, unsigned mask, unsigned long long key, int i)
{
    unsigned slot = hash64(key) & mask;
    while (true) {
        unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(keys + slot);
        if (k == EMPTY) {
            k = atomicCAS(keys + slot, EMPTY, key);
            if (k == EMPTY) k = key;
        }
        if (k == key) {
            if (*reinterpret_cast<volatile int *>(vals + slot) > i) atomicMin(vals + slot, i);
            return slot;
        }
        slot = (slot + 1) & mask;
    }
}

__global__ void __launch_bounds__(256) voxel_claim_kernel(const VoxArgs a)
{
    const int f = blockIdx.y;
    const long long beg = a.offsets[f], n = a.offsets[f + 1] - beg;
    const float4 *p4 = reinterpret_cast<const float4 *>(a.pts) + beg;
    unsigned long long *keys = a.keys + (size_t)f * 4 * a.capT;
    int *vals = a.vals + (size_t)f * 4 * a.capT;
    const unsigned mask = a.capT - 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float4 p = p4[i];
        unsigned s0 = NOSLOT, s1 = NOSLOT, s2 = NOSLOT, sb = NOSLOT;
        VoxelOfPoint v;
        const int st = voxel_of_point(p.x, p.y, p.z, v);
        if (st < 0) {
            if (a.status) atomicAdd(a.status + f, 1);                // the reference raises IndexError
        } else if (st > 0) {
            sb = claim(keys + 3 * (size_t)a.capT, vals + 3 * (size_t)a.capT, mask, vkey(v.b[0], v.b[1], v.b[2]), (int)i);
            s0 = claim(keys, vals, mask, vkey(v.g0[0], v.g0[1], v.g0[2]), (int)i);
            s1 = claim(keys + a.capT, vals + a.capT, mask, vkey(v.g1[0], v.g1[1], v.g1[2]), (int)i);
            s2 = claim(keys + 2 * (size_t)a.capT, vals + 2 * (size_t)a.capT, mask, vkey(v.g2[0], v.g2[1], v.g2[2]), (int)i);
        }
        const long long g = beg + i;
        a.slots[g] = s0;
        a.slots[a.total_points + g] = s1;
        a.slots[2 * a.total_points + g] = s2;
        a.slots[3 * a.total_points + g] = sb;
    }
}

__device__ __forceinline__ void store3(short *dst, unsigned long long key)
{
    dst[0] = (short)(key & 0xFFFF);
    dst[1] = (short)((key >> 16) & 0xFFFF);
    dst[2] = (short)((key >> 32) & 0xFFFF);
}

// One CTA per (table t, frame f): ordered compaction of the points that own their key.
constexpr int CP_THREADS = 1024, CP_PER = 4;
__global__ void __launch_bounds__(CP_THREADS) voxel_compact_kernel(const VoxArgs a)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int t = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long beg = a.offsets[f];
    const int n = (int)(a.offsets[f + 1] - beg);
    const unsigned *slots = a.slots + (size_t)t * a.total_points + beg;
    const unsigned *bslots = a.slots + (size_t)3 * a.total_points + beg;
    const unsigned long long *keys = a.keys + ((size_t)f * 4 + t) * a.capT;
    const int *vals = a.vals + ((size_t)f * 4 + t) * a.capT;
    const size_t fcap = (size_t)f * a.cap;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int tile = 0; tile < n; tile += CP_THREADS * CP_PER) {
        const int i0 = tile + tid * CP_PER;
        unsigned sl[CP_PER];
        bool own[CP_PER];
        int c = 0;
#pragma unroll
        for (int k = 0; k < CP_PER; ++k) sl[k] = (i0 + k < n) ? slots[i0 + k] : NOSLOT;
#pragma unroll
        for (int k = 0; k < CP_PER; ++k) {
            own[k] = sl[k] != NOSLOT && vals[sl[k]] == i0 + k;
            c += own[k];
        }
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
            s_warp[lane] = wi - w;  // exclusive
        }
        __syncthreads();
        int rank = s_base + s_warp[warp] + inc - c;
#pragma unroll
        for (int k = 0; k < CP_PER; ++k) {
            if (!own[k]) continue;
            if (rank < a.cap) {
                if (t == 0) {
                    a.w0_slot[fcap + rank] = sl[k];
                    a.w0_bslot[fcap + rank] = bslots[i0 + k];
                } else if (t == 3) {
                    a.brank[(size_t)f * a.capT + sl[k]] = (unsigned)rank;
                    if (a.blocks) store3(a.blocks + (fcap + rank) * 3, keys[sl[k]]);
                } else {
                    store3(a.vox + (((size_t)f * 3 + t) * a.cap + rank) * 3, keys[sl[k]]);
                }
            }
            ++rank;
        }
        __syncthreads();
        if (tid == CP_THREADS - 1) s_base = rank;
        __syncthreads();
    }
    if (tid == 0) a.counts[f * 4 + t] = s_base;
}

// per scale-0 winner: look up its block rank, count voxels per block
__global__ void __launch_bounds__(256) voxel_blockcount_kernel(const VoxArgs a)
{
    const int f = blockIdx.y;
    const int n0 = min(a.counts[f * 4 + 0], a.cap);
    const size_t fcap = (size_t)f * a.cap;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n0; j += gridDim.x * blockDim.x) {
        const unsigned b = a.brank[(size_t)f * a.capT + a.w0_bslot[fcap + j]];
        a.w0_blk[fcap + j] = b;
        atomicAdd(a.bcount + fcap + b, 1);
    }
}

// one CTA per frame: cnt = exclusive scan of bcount (cntVoxelsLength, Voxel.py:161-163)
__global__ void __launch_bounds__(1024) voxel_blockscan_kernel(const VoxArgs a)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = min(a.counts[f * 4 + 3], a.cap);
    const int *bc = a.bcount + (size_t)f * a.cap;
    int *cnt = a.cnt + (size_t)f * (a.cap + 1);
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int tile = 0; tile < nb; tile += 1024) {
        const int i = tile + tid;
        const int c = i < nb ? bc[i] : 0;
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
        if (i < nb) cnt[i] = excl;
        __syncthreads();
        if (tid == 1023) s_base = excl + c;
        __syncthreads();
    }
    if (tid == 0) cnt[nb] = s_base;
}

__global__ void __launch_bounds__(256) voxel_segclaim_kernel(const VoxArgs a)
{
    const int f = blockIdx.y;
    const int n0 = min(a.counts[f * 4 + 0], a.cap);
    const size_t fcap = (size_t)f * a.cap;
    const int *cnt = a.cnt + (size_t)f * (a.cap + 1);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n0; j += gridDim.x * blockDim.x) {
        const unsigned b = a.w0_blk[fcap + j];
        const int pos = cnt[b] + atomicAdd(a.bfill + fcap + b, 1);
        a.seg[fcap + pos] = (unsigned)j;
        a.seg_blk[fcap + pos] = b;
    }
}

// stable order inside each block segment: rank = how many winners of the same block came earlier
__global__ void __launch_bounds__(256) voxel_segrank_kernel(const VoxArgs a)
{
    const int f = blockIdx.y;
    const int n0 = min(a.counts[f * 4 + 0], a.cap);
    const size_t fcap = (size_t)f * a.cap;
    const int *cnt = a.cnt + (size_t)f * (a.cap + 1);
    const unsigned long long *keys0 = a.keys + (size_t)f * 4 * a.capT;
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n0; pos += gridDim.x * blockDim.x) {
        const unsigned b = a.seg_blk[fcap + pos], j = a.seg[fcap + pos];
        const int s0 = cnt[b], s1 = cnt[b + 1];
        int rank = 0;
        for (int q = s0; q < s1; ++q) rank += a.seg[fcap + q] < j;
        const unsigned long long key = keys0[a.w0_slot[fcap + j]];
        const size_t o = (size_t)s0 + rank;
        store3(a.vox + ((size_t)f * 3 * a.cap + o) * 3, key);
        if (a.local0) {
            short *l = a.local0 + (fcap + o) * 3;
            l[0] = (short)(key & 63);
            l[1] = (short)((key >> 16) & 63);
            l[2] = (short)((key >> 32) & 63);
        }
    }
}

int upload_offsets(caelo_ctx *ctx, const int64_t *host, int n, long long *dev, cudaStream_t st)
{
    void *h = nullptr;
    cudaEvent_t ev;
    int rc = caelo_stage_acquire(ctx, (size_t)n * 8, &h, &ev);
    if (rc) return rc;
    memcpy(h, host, (size_t)n * 8);
    CAELO_CUDA(ctx, caelo_stage_copy_async(dev, h, (size_t)n * 8, st));
    CAELO_CUDA(ctx, cudaEventRecord(ev, st));
    return CAELO_OK;
}

}  // namespace

extern "C" int caelo_project_ring(caelo_ctx *ctx, const float *pts, const int64_t *pts_offsets, int F, float *ring5,
                                  int32_t *counter_i32, float *ring3, int8_t *counter_i8, int32_t *status, void *stream)
{
    if (!ctx || !pts || !pts_offsets || F <= 0) return CAELO_ERR_ARG;
    if (!ring5 && !counter_i32 && !ring3 && !counter_i8) return CAELO_ERR_ARG;
    long long maxn = 0;
    for (int f = 0; f < F; ++f) {
        long long n = pts_offsets[f + 1] - pts_offsets[f];
        if (n < 0 || n > 0x7FFFFFFFLL) return CAELO_ERR_ARG;
        if (n > maxn) maxn = n;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npx = (size_t)F * RING_H * RING_W;
    const size_t head = ((size_t)(F + 1) * 8 + 255) / 256 * 256;
    int rc = caelo_reserve(ctx, ctx->scan_ws, head + npx * 8);
    if (rc) return rc;
    char *base = reinterpret_cast<char *>(ctx->scan_ws.ptr);
    long long *d_off = reinterpret_cast<long long *>(base);
    if ((rc = upload_offsets(ctx, pts_offsets, F + 1, d_off, st))) return rc;
    RingArgs a;
    a.pts = pts; a.offsets = d_off;
    a.winner = reinterpret_cast<int *>(base + head);
    a.hits = a.winner + npx;
    a.status = status;
    // SphericalRing.py:34-57, derived exactly as the reference derives them
    const double d2r = 3.14159265358979323846 / 180;
    const double vdown = -24.8 * d2r, vup = 2.0 * d2r;
    a.az_res = 0.20 * d2r;
    a.v_res = (vup - vdown) / (64 - 1);
    a.v_off = -vdown / a.v_res;
    a.ring5 = ring5; a.counter_i32 = counter_i32; a.ring3 = ring3; a.counter_i8 = reinterpret_cast<signed char *>(counter_i8);
    a.F = F;
    CAELO_CUDA(ctx, caelo_fill_async(a.winner, 0xFF, npx * 4, st));
    CAELO_CUDA(ctx, caelo_fill_async(a.hits, 0, npx * 4, st));
    if (status) CAELO_CUDA(ctx, caelo_fill_async(status, 0, (size_t)F * 4, st));
    int bx = (int)((maxn + 255) / 256);
    if (bx > 128) bx = 128;
    if (bx < 1) bx = 1;
    { ProfScope ps_(ctx, "ring_claim_kernel", st); ring_claim_kernel<<<dim3(bx, F), 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "ring_fill_kernel", st); ring_fill_kernel<<<dim3(64, F), 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}

extern "C" int caelo_voxelize(caelo_ctx *ctx, const float *pts, const int64_t *pts_offsets, int F, int cap, int16_t *vox,
                              int32_t *counts, int16_t *local0, int16_t *blocks, int32_t *cnt, int32_t *status,
                              void *stream)
{
    if (!ctx || !pts || !pts_offsets || !vox || !counts || F <= 0 || cap <= 0) return CAELO_ERR_ARG;
    long long maxn = 0;
    for (int f = 0; f < F; ++f) {
        long long n = pts_offsets[f + 1] - pts_offsets[f];
        if (n < 0 || n > cap) return CAELO_ERR_ARG;   // a list can hold one voxel per point
        if (n > maxn) maxn = n;
    }
    const long long total = pts_offsets[F];   // rows of pts addressed by the offsets
    cudaStream_t st = (cudaStream_t)stream;
    unsigned capT = 1024;
    while (capT < 2 * (unsigned long long)maxn + 2) capT <<= 1;
    // scratch layout
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_off = take((size_t)(F + 1) * 8);
    const size_t o_keys = take((size_t)F * 4 * capT * 8);
    const size_t o_vals = take((size_t)F * 4 * capT * 4);
    const size_t o_brank = take((size_t)F * capT * 4);
    const size_t o_slots = take((size_t)4 * total * 4);
    const size_t o_w = take((size_t)5 * F * cap * 4);
    const size_t o_bc = take((size_t)2 * F * cap * 4);
    const size_t o_cnt = take((size_t)F * (cap + 1) * 4);
    int rc = caelo_reserve(ctx, ctx->scan_ws, off);
    if (rc) return rc;
    char *base = reinterpret_cast<char *>(ctx->scan_ws.ptr);
    VoxArgs a;
    a.pts = pts;
    a.offsets = reinterpret_cast<long long *>(base + o_off);
    if ((rc = upload_offsets(ctx, pts_offsets, F + 1, const_cast<long long *>(a.offsets), st))) return rc;
    a.keys = reinterpret_cast<unsigned long long *>(base + o_keys);
    a.vals = reinterpret_cast<int *>(base + o_vals);
    a.brank = reinterpret_cast<unsigned *>(base + o_brank);
    a.slots = reinterpret_cast<unsigned *>(base + o_slots);
    a.total_points = total;
    a.capT = capT;
    unsigned *w = reinterpret_cast<unsigned *>(base + o_w);
    const size_t fc = (size_t)F * cap;
    a.w0_slot = w; a.w0_bslot = w + fc; a.w0_blk = w + 2 * fc; a.seg = w + 3 * fc; a.seg_blk = w + 4 * fc;
    a.bcount = reinterpret_cast<int *>(base + o_bc);
    a.bfill = a.bcount + fc;
    a.vox = vox; a.local0 = local0; a.blocks = blocks;
    a.cnt = cnt ? cnt : reinterpret_cast<int *>(base + o_cnt);
    a.counts = counts; a.status = status; a.cap = cap; a.F = F;
    CAELO_CUDA(ctx, caelo_fill_async(a.keys, 0xFF, (size_t)F * 4 * capT * 8, st));
    CAELO_CUDA(ctx, caelo_fill_async(a.vals, 0x7F, (size_t)F * 4 * capT * 4, st));
    CAELO_CUDA(ctx, caelo_fill_async(a.bcount, 0, 2 * fc * 4, st));
    if (status) CAELO_CUDA(ctx, caelo_fill_async(status, 0, (size_t)F * 4, st));
    int bx = (int)((maxn + 255) / 256);
    if (bx > 128) bx = 128;
    if (bx < 1) bx = 1;
    { ProfScope ps_(ctx, "voxel_claim_kernel", st); voxel_claim_kernel<<<dim3(bx, F), 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "voxel_compact_kernel", st); voxel_compact_kernel<<<dim3(4, F), CP_THREADS, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "voxel_blockcount_kernel", st); voxel_blockcount_kernel<<<dim3(bx, F), 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "voxel_blockscan_kernel", st); voxel_blockscan_kernel<<<F, 1024, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "voxel_segclaim_kernel", st); voxel_segclaim_kernel<<<dim3(bx, F), 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    { ProfScope ps_(ctx, "voxel_segrank_kernel", st); voxel_segrank_kernel<<<dim3(bx, F), 256, 0, st>>>(a); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}
