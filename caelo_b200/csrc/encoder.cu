// encoder.cu — a3: PatchEncoder.predict (reference Match.py:130-135; graph AE4VoxelPatch.py:189-197
// with the activations of the shipped EncoderModel4VoxelPatch.h5 = tanh on every layer).
//
//   conv3d 3^3 1->8 tanh, maxpool 2 | conv3d 8->16 tanh, maxpool 2 | conv3d 16->32 tanh |
//   flatten (x,y,z,c) | dense 2048->200 tanh | dense 200->20 tanh
//
// Input is the 512-byte bit-packed occupancy patch produced by patches.cu (values are exactly
// {0,1}); max-pool commutes with the monotonic tanh, so tanh is applied after pooling.
//
//   conv12_pair_kernel conv1 (CUDA cores: exact sums of selected weights from pattern tables, max-pool, tanh, split fp16) +
//                      conv2 (tcgen05 implicit GEMM) for TWO patches per MMA (M = 128) with the dx taps folded into N;
//                      one persistent CTA per SM: 16 producer warps, issuer warp, 8 epilogue warps -> act2 (the default)
//   conv12_tc_kernel   the round-1 form (one patch per M = 64 MMA, 2 CTAs per SM), kept behind CAELO_CONV12_PAIR=0
//   conv3_oct_kernel   conv3 for EIGHT patches per MMA (M = 128, N = 192 with dx folded in), slab-wise operand ring of four
//                      stages, 8 producer warps, issuer warp, 16 epilogue warps -> act3 as split fp16, tile-major (the default)
//   conv3_tc_kernel    the round-1 form (one patch per M = 64 MMA, 9 shifted compact copies), kept behind CAELO_CONV3_OCT=0
//   dense_tc_kernel    256 patches x 208 outputs x K = 2048 per CTA, tile-major operands by 1-D bulk copies (cp.async.bulk +
//                      mbarrier complete_tx), 3-stage mbarrier pipeline, epilogue tanh + dense2 (200 -> 20) + tanh in registers
// Every product runs as split fp16 (x = hi + lo) with fp32 accumulation in TMEM; descriptors within 1e-4 (measured 8e-6).
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace {

// tanh(x) = sign(x) (1 - 2/(exp(2|x|)+1)) on the SFU (ex2.approx + rcp): |err| < 3e-7 absolute
__device__ __forceinline__ float fast_tanh(float x)
{
    float ax = fabsf(x);
    float e = __expf(2.0f * ax);
    float t = 1.0f - __fdividef(2.0f, e + 1.0f);
    return copysignf(t, x);
}

// ---- conv1 (CUDA cores) + conv2 (tcgen05) --------------------------------------------------
constexpr int TC_WORKERS = 256;                       // warps 0-7: produce operands (conv1)
constexpr int TC_ISSUER = 8;                          // warp 8: its elected lane issues the MMAs (absorbs the tensor-queue back-pressure)
constexpr int TC_EPI_WARPS = 4;                       // warps 9-12: drain accumulators (one per TMEM lane quarter)
constexpr int TC_THREADS = TC_WORKERS + 32 + TC_EPI_WARPS * 32;
constexpr int A_VOL_BYTES = 10 * 10 * 10 * 16;        // padded 8^3 volume, 8 ch fp16 per position
constexpr int W2_BYTES = 28 * 512;                    // 28 tap chunks x 32 rows x 16 B
// dynamic smem layout (bytes)
constexpr int SM_A = 0;                               // [buf 2][hi,lo][A_VOL_BYTES]
constexpr int SM_W2 = SM_A + 4 * A_VOL_BYTES;         // 64000
constexpr int T1_PAT = 48;                            // a partial-sum row (32 B) every 48 B: the 8 rows of a nibble hit 8 distinct 16-B bank groups
constexpr int T1_ROW = 8 * T1_PAT;                    // one nibble's 8 partial-sum rows
constexpr int SM_T1 = SM_W2 + W2_BYTES;               // conv1 partial sums [9 (dx,dy)][8 dz-patterns][8 ch] f32
constexpr int SM_B12 = SM_T1 + 9 * T1_ROW;            // floats: b1[8] b2[16]
constexpr int SM_BG = SM_B12 + (8 + 16) * 4;          // tanh(b1) as fp16 hi (16 B) + lo (16 B)
constexpr int PK_PITCH = 20;                          // staged occupancy rows: [18 (x halo)][20 (y halo, padded)] u16
constexpr int PK_BYTES = 18 * PK_PITCH * 2;           // 720
constexpr int SM_PK = SM_BG + 32;                     // [2][PK_BYTES]
constexpr int SM_LWIN = SM_PK + 2 * PK_BYTES;         // non-empty cell list: windows [512] u64
constexpr int SM_LCELL = SM_LWIN + 512 * 8;           //                      padded cell index [2 buffers][512] u16
constexpr int SM_LCNT = SM_LCELL + 2 * 512 * 2;       //                      count [2] (double-buffered)
constexpr int SM_XS = SM_LCNT + 16;                   // per patch (ring of 4): 8 flags "x-slice holds a non-empty cell"
constexpr int SM_BGP = SM_XS + 32;                    // conv2+pool+tanh of an all-background neighbourhood [27 classes][16]
constexpr int SM_BAR = SM_BGP + 27 * 16 * 4;          // 6 mbarriers + tmem base
constexpr int TC_SMEM = SM_BAR + 64;
static_assert(SM_T1 % 16 == 0 && SM_LWIN % 8 == 0 && SM_XS % 8 == 0 && SM_BGP % 8 == 0 && SM_BAR % 8 == 0, "smem alignment");

struct Conv12Args {
    const unsigned *packed;  // [P,128]
    const float *k1, *b1;    // (27,8), (8)
    const float *k2, *b2;    // (27,8,16), (16)
    const float *tables;     // prep_conv12_tables_kernel: T1 [9][8][8] then background [27][16]
    float *act2;             // [P,64,16] fp32
    int P;
    int K3;                  // frame mode: 3*K (patch order [F][3][K]) — CTAs then walk the patches scale-interleaved; 0 = as stored
    int skip_bg;             // 1: x-slice pairs whose whole conv2 neighbourhood is background skip their MMAs
    int dbg;                 // measurement only (CAELO_CONV12_DBG): bit 0 = no MMAs, bit 1 = no conv1 pass 2 (results are wrong)
    long long *timeline;     // debug: [gridDim.x][64][16] clock64 stamps, or null
};

__device__ __forceinline__ constexpr int tap_off(int t)  // byte offset of tap t inside the padded volume
{
    return (((t / 9) * 10 + (t / 3) % 3) * 10 + t % 3) * 16;
}

// conv1 + maxpool + tanh for one patch -> A_hi / A_lo interior, in two passes so that the work is
// balanced over the 256 worker threads however the occupied voxels cluster.  The volumes are kept
// "clean": every interior cell holds tanh(b1) (conv1 of an empty neighbourhood) unless a patch wrote it,
// and the cells a patch wrote are restored before the buffer is reused — so empty cells cost nothing.
//   pass 1  thread = one (px,py) column, two pz cells: the 4x4 occupancy rows around the column are read
//           from the halo-padded staging copy (8 aligned 32-bit loads); cells with a non-empty 4x4x4
//           window are appended (cell, window) to a list in shared memory;
//   pass 2  listed cells are dealt 4 per warp, EIGHT lanes per cell — one per pooled sub-position
//           (sx,sy,sz).  A sub-position's 27 neighbourhood bits are 9 (dx,dy) groups of 3 dz bits; each
//           group indexes a table of precomputed partial weight sums (9 x 8 patterns x 8 channels), so the
//           conv is 9 table rows added up: no data-dependent loop, every load independent.  Max over the
//           sub-positions by three channel-halving exchanges, tanh, split to fp16 hi/lo — each lane
//           finishes one channel.
//   The operand cell of pooled position (px,py,pz) is  px*SX + py*SY + pz + c0  (16 B each): the one-patch kernel keeps a
//   10^3 volume (SX = 100, SY = 10, c0 = 111), the pair kernel an [x 8][y 10][patch 2][z 10] volume (SX = 200, SY = 20,
//   c0 = 21 + 10*patch).  `barid` = the named barrier of the 256 threads that work on this patch.
template <int SX, int SY>
__device__ __forceinline__ void conv1_to_smem(const unsigned short *rows, const unsigned char *t1, const float *b1s,
                                              unsigned char *a_hi, unsigned char *a_lo,
                                              unsigned long long *lwin, unsigned short *lcell, int *lcount,
                                              unsigned char *xs_any, int tid, int &n_listed, long long *tl,
                                              int c0 = 111, int barid = 1)
{
    const int lane = tid & 31;
    {
        const int col = tid >> 2, zq = tid & 3;
        const int px = col >> 3, py = col & 7;
        unsigned r[16];
        unsigned any = 0, cells = 0;
        // staged row (x,y) lives at [(x+1)*PK_PITCH + (y+1)]; the column needs x = 2px-1.., y = 2py-1.. (4 each)
        const unsigned *rp = reinterpret_cast<const unsigned *>(rows + (2 * px) * PK_PITCH + 2 * py);
#pragma unroll
        for (int ix = 0; ix < 4; ++ix) {
            const unsigned v01 = rp[ix * (PK_PITCH / 2)], v23 = rp[ix * (PK_PITCH / 2) + 1];
            r[ix * 4 + 0] = (v01 & 0xFFFFu) << 1;  // bit z+1 <-> voxel z
            r[ix * 4 + 1] = (v01 >> 16) << 1;
            r[ix * 4 + 2] = (v23 & 0xFFFFu) << 1;
            r[ix * 4 + 3] = (v23 >> 16) << 1;
            any |= v01 | v23;
        }
        any = (any | (any >> 16)) & 0xFFFFu;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int pz = 2 * zq + half;
            unsigned wlo = 0u, whi = 0u;
            if (((any << 1) >> (2 * pz)) & 0xFu) {   // some row has a voxel at z = 2pz-1 .. 2pz+2
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    wlo |= ((r[i] >> (2 * pz)) & 0xFu) << (i * 4);
                    whi |= ((r[8 + i] >> (2 * pz)) & 0xFu) << (i * 4);
                }
            }
            const bool nz = (wlo | whi) != 0u;
            const unsigned m = __ballot_sync(0xffffffffu, nz);
            cells |= m;
            if (m) {
                int base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(lcount, __popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (nz) {
                    const int k = base + __popc(m & ((1u << lane) - 1u));
                    lwin[k] = ((unsigned long long)whi << 32) | wlo;
                    lcell[k] = (unsigned short)(px * SX + py * SY + pz + c0);
                }
            }
        }
        if (lane == 0) xs_any[tid >> 5] = cells != 0u;   // warp w owns the cells of x-slice px = w
    }
    asm volatile("bar.sync %0, 256;" ::"r"(barid) : "memory");
    if (tl && tid == 0) tl[7] = clock64();
    // pass 2: eight consecutive lanes share a listed cell, lane s = (sx,sy,sz); a warp takes 4 cells per round
    // (warp-uniform trip count: the shuffles below need every lane)
    const int n = *lcount;
    n_listed = n;
    const int sub = tid & 7, sx = sub >> 2, sy = (sub >> 1) & 1, sz = sub & 1;
    for (int k0 = (tid >> 5) * 4; k0 < n; k0 += (TC_WORKERS / 32) * 4) {
        const int k = k0 + (lane >> 3);
        const bool valid = k < n;
        const unsigned long long win = valid ? lwin[k] : 0ull;
        // 3x3 nibbles (dx,dy) of the window around (sx,sy): nibble j = dx*3+dy at bits 4j..4j+3; dz = 0..2 above sz
        const unsigned long long w2 = win >> (16 * sx + 4 * sy + sz);
        const unsigned g0 = (unsigned)w2 & 0x777u, g1 = (unsigned)(w2 >> 16) & 0x777u, g2 = (unsigned)(w2 >> 32) & 0x777u;
        float2 acc2[4];   // channels (0,1) (2,3) (4,5) (6,7): packed f32x2 adds, one instruction per two channels
        {
            const float4 c0 = *reinterpret_cast<const float4 *>(b1s), c1 = *reinterpret_cast<const float4 *>(b1s + 4);
            acc2[0] = make_float2(c0.x, c0.y); acc2[1] = make_float2(c0.z, c0.w);
            acc2[2] = make_float2(c1.x, c1.y); acc2[3] = make_float2(c1.z, c1.w);
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const unsigned gj = (j < 3) ? g0 : ((j < 6) ? g1 : g2);
            const unsigned pat = (gj >> (4 * (j % 3))) & 7u;
            // the patches are sparse: most (dx,dy) groups of a listed sub-position are empty, and the empty pattern's
            // row is all zeros (x + 0 = x exactly) — lanes with pat == 0 skip the two 16-byte loads, which cuts the
            // shared-memory wavefronts the tensor core's operand fetch competes with (ncu round 2: the LSU took 60 %
            // of the shared-memory pipe next to the tensor core's 74 %, three quarters of it these table rows)
            if (pat) {
                const float4 *row = reinterpret_cast<const float4 *>(t1 + j * T1_ROW + pat * T1_PAT);
                const float4 w0 = row[0], w1 = row[1];
                acc2[0] = __fadd2_rn(acc2[0], make_float2(w0.x, w0.y));
                acc2[1] = __fadd2_rn(acc2[1], make_float2(w0.z, w0.w));
                acc2[2] = __fadd2_rn(acc2[2], make_float2(w1.x, w1.y));
                acc2[3] = __fadd2_rn(acc2[3], make_float2(w1.z, w1.w));
            }
        }
        const float acc[8] = {acc2[0].x, acc2[0].y, acc2[1].x, acc2[1].y, acc2[2].x, acc2[2].y, acc2[3].x, acc2[3].y};
        // max over the eight sub-position lanes, halving the channel set a lane carries at each exchange:
        // lane `sub` ends with channel 4*sx + 2*sy + sz
        float h4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float recv = __shfl_xor_sync(0xffffffffu, sx ? acc[c] : acc[4 + c], 4);
            h4[c] = fmaxf(sx ? acc[4 + c] : acc[c], recv);
        }
        float h2[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float recv = __shfl_xor_sync(0xffffffffu, sy ? h4[c] : h4[2 + c], 2);
            h2[c] = fmaxf(sy ? h4[2 + c] : h4[c], recv);
        }
        const float recv = __shfl_xor_sync(0xffffffffu, sz ? h2[0] : h2[1], 1);
        const float o = fmaxf(sz ? h2[1] : h2[0], recv);
        if (valid) {
            const int pi = lcell[k];
            __half h, l;
            umma::split_f16(fast_tanh(o), h, l);
            *reinterpret_cast<__half *>(a_hi + pi * 16 + sub * 2) = h;
            *reinterpret_cast<__half *>(a_lo + pi * 16 + sub * 2) = l;
        }
    }
    if (tl && tid == 0) tl[8] = clock64();
}

__global__ void __launch_bounds__(TC_THREADS, 2) conv12_tc_kernel(const Conv12Args a)   // 416 threads x 2 CTAs: <= 78 registers
{
    extern __shared__ __align__(128) unsigned char sm[];
    float *b1s = reinterpret_cast<float *>(sm + SM_B12), *b2s = b1s + 8;
    const uint4 *bg = reinterpret_cast<const uint4 *>(sm + SM_BG);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + SM_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + SM_BAR + 48);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler

    // ---- one-time setup: zero the operand volumes and the staged rows (halos stay zero), stage weights/tables ----
    for (int i = tid; i < SM_T1 / 16; i += TC_THREADS) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 2 * PK_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4 *>(sm + SM_PK)[i] = make_uint4(0, 0, 0, 0);
    for (int e = tid; e < 9 * 8 * 8; e += TC_THREADS) {       // T1 row (j,pat) at j*T1_ROW + pat*T1_PAT
        const int j = e >> 6, pat = (e >> 3) & 7, c = e & 7;
        *reinterpret_cast<float *>(sm + SM_T1 + j * T1_ROW + pat * T1_PAT + c * 4) = a.tables[e];
    }
    for (int e = tid; e < 27 * 16; e += TC_THREADS) reinterpret_cast<float *>(sm + SM_BGP)[e] = a.tables[576 + e];
    if (tid < 8) {
        b1s[tid] = a.b1[tid];
        __half h, l;
        umma::split_f16(tanhf(a.b1[tid]), h, l);
        reinterpret_cast<__half *>(sm + SM_BG)[tid] = h;
        reinterpret_cast<__half *>(sm + SM_BG + 16)[tid] = l;
    }
    if (tid < 16) b2s[tid] = a.b2[tid];
    __syncthreads();
    // every interior cell of the four operand volumes starts as the background value tanh(b1)
    for (int e = tid; e < 4 * 512; e += TC_THREADS) {
        const int vol = e >> 9, c = e & 511;
        const int pi = (((c >> 6) + 1) * 10 + ((c >> 3) & 7) + 1) * 10 + (c & 7) + 1;
        *reinterpret_cast<uint4 *>(sm + SM_A + vol * A_VOL_BYTES + pi * 16) = bg[vol & 1];
    }
    // B operand: row n (0..15 = W_hi, 16..31 = W_lo of out-channel n%16), k = tap*8 + ci; chunk = tap
    for (int e = tid; e < 27 * 8 * 16; e += TC_THREADS) {
        int t = e / 128, ci = (e / 16) % 8, co = e % 16;
        __half h, l;
        umma::split_f16(a.k2[e], h, l);
        unsigned char *w = sm + SM_W2 + t * 512;
        *reinterpret_cast<__half *>(w + (co / 8) * 128 + (co % 8) * 16 + ci * 2) = h;
        *reinterpret_cast<__half *>(w + ((16 + co) / 8) * 128 + (co % 8) * 16 + ci * 2) = l;
    }
    if (warp == TC_ISSUER) umma::tmem_alloc(tmem_slot, 256);
    uint64_t *full = mbar, *tfull = mbar + 2, *tempty = mbar + 4;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&full[b], TC_WORKERS);    // operands of buffer b written (every worker arrives)
            umma::mbar_init(&tfull[b], 1);            // MMAs into TMEM[b] complete (tcgen05.commit)
            umma::mbar_init(&tempty[b], TC_EPI_WARPS * 32);  // TMEM[b] drained by the epilogue warps
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
    // Patches of the three scales differ a lot in conv1 work (2 / 46 / 127 voxels on average) but not in
    // conv2 work: in frame mode consecutive patches of a CTA (and the two CTAs of an SM) cycle through the
    // scales, so CUDA-core-heavy and tensor-heavy patches overlap instead of arriving in phases.
    auto patch_of = [&](int i) {
        const int v = (int)blockIdx.x + i * (int)gridDim.x;
        if (a.K3 == 0) return v;
        const int f = v / a.K3, r = v - f * a.K3, K = a.K3 / 3;
        return f * a.K3 + (r % 3) * K + r / 3;
    };
    auto stamp = [&](int i, int slot) {
        if (a.timeline && i < 64) a.timeline[((size_t)blockIdx.x * 64 + i) * 16 + slot] = clock64();
    };

    // x-slice pairs (= one pooled x) of buffer b that need conv2: some cell of slices 2p-1 .. 2p+2 is non-empty
    auto active_pairs = [&](int b) -> unsigned {
        if (!a.skip_bg) return 0xFu;
        const unsigned long long f = *reinterpret_cast<const unsigned long long *>(sm + SM_XS + 8 * b);   // b = patch ordinal & 3
        unsigned m = 0;  // bit xs: slice xs holds a non-empty cell
#pragma unroll
        for (int xs = 0; xs < 8; ++xs) m |= (unsigned)((f >> (8 * xs)) & 1ull) << xs;
        const unsigned near = m | (m << 1) | (m >> 1);  // slice xs has a non-empty cell within +-1
        unsigned act = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) act |= ((near >> (2 * p)) & 3u) ? (1u << p) : 0u;
        return act;
    };

    if (warp == TC_ISSUER) {
        // ===== MMA issuer: waits for operands + a free accumulator, queues 224 MMAs, commits =====
        const uint32_t idesc32 = umma::idesc_f16_f32(64, 32), idesc16 = umma::idesc_f16_f32(64, 16);
        for (int j = 0; j < n_my; ++j) {
            const int b = j & 1, k = j >> 1;
            umma::mbar_wait(&full[b], (uint32_t)(k & 1));
            if (k >= 1) umma::mbar_wait(&tempty[b], (uint32_t)((k - 1) & 1));
            umma::fence_after_thread_sync();
            if (lane == 0) stamp(j, 5);
            const unsigned act = active_pairs(j & 3);
            if (umma::elect_one()) {
#pragma unroll 1
                for (int xs = 0; xs < 8; ++xs) {  // x-slice: M = 64 positions (y,z)
                    if (!((act >> (xs >> 1)) & 1u)) continue;
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
    } else if (warp < TC_ISSUER) {
        // ===== producers (8 warps): conv1 of patch i into operand buffer i&1, as soon as MMA(i-2) has released it =====
        unsigned long long *lwin = reinterpret_cast<unsigned long long *>(sm + SM_LWIN);
        unsigned short *lcell = reinterpret_cast<unsigned short *>(sm + SM_LCELL);
        int *lcnt = reinterpret_cast<int *>(sm + SM_LCNT);
        unsigned pk_next = 0u;  // packed word of the next patch to produce, fetched one iteration ahead
        int p_next = 0;         // its patch index
        int n_dirty0 = 0, n_dirty1 = 0;  // cells of each buffer that the patch before wrote (to be restored to the background)
        auto fetch = [&](int i) {
            if (i < n_my) {
                p_next = patch_of(i);
                if (tid < 128) pk_next = __ldg(a.packed + (size_t)p_next * 128 + tid);
            }
        };
        fetch(0);
        auto produce = [&](int i) {
            const int b = i & 1;
            unsigned char *a_hi = sm + SM_A + (2 * b) * A_VOL_BYTES, *a_lo = a_hi + A_VOL_BYTES;
            unsigned short *rows = reinterpret_cast<unsigned short *>(sm + SM_PK + b * PK_BYTES);
            unsigned short *lc = lcell + b * 512;
            // restore the cells the previous patch of this buffer wrote (its MMAs are complete: tfull was waited on)
            const int nd = b ? n_dirty1 : n_dirty0;
            for (int k = tid; k < nd; k += TC_WORKERS) {
                const int pi = lc[k];
                *reinterpret_cast<uint4 *>(a_hi + pi * 16) = bg[0];
                *reinterpret_cast<uint4 *>(a_lo + pi * 16) = bg[1];
            }
            if (tid < 128) {  // word tid = occupancy rows (x, y) and (x, y+1), x = tid/8, y = 2*(tid%8)
                unsigned short *dst = rows + ((tid >> 3) + 1) * PK_PITCH + 2 * (tid & 7) + 1;
                dst[0] = (unsigned short)(pk_next & 0xFFFFu);
                dst[1] = (unsigned short)(pk_next >> 16);
            }
            fetch(i + 1);
            if (tid == 128) lcnt[b] = 0;  // the other counter was read before the previous patch's second barrier
            asm volatile("bar.sync 1, 256;" ::: "memory");
            long long *tl = (a.timeline && i < 64) ? a.timeline + ((size_t)blockIdx.x * 64 + i) * 16 : nullptr;
            if (tl && tid == 0) tl[2] = clock64();
            int n_listed;
            conv1_to_smem<100, 10>(rows, sm + SM_T1, b1s, a_hi, a_lo, lwin, lc, lcnt + b, sm + SM_XS + 8 * (i & 3), tid, n_listed, tl);
            if (b) n_dirty1 = n_listed; else n_dirty0 = n_listed;
            umma::fence_proxy_async();
            umma::mbar_arrive(&full[b]);
        };
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            if (tid == 0) stamp(i, 0);
            if (i >= 2) {   // MMA(i-2) must have finished reading operand buffer b
                umma::mbar_wait(&tfull[b], (uint32_t)(((i - 2) >> 1) & 1));
                umma::fence_after_thread_sync();
            }
            produce(i);
            if (tid == 0) stamp(i, 1);
        }
    } else {
        // ===== epilogue (4 warps, one per TMEM lane quarter): drain patch i while conv1(i+1) and MMA(i+1) run =====
        const int q = warp & 3;
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            umma::mbar_wait(&tfull[b], (uint32_t)((i >> 1) & 1));
            umma::fence_after_thread_sync();
            if (warp == TC_ISSUER + 1 && lane == 0) stamp(i, 3);
            float *out = a.act2 + (size_t)patch_of(i) * 1024;
            const unsigned act = active_pairs(i & 3);
#pragma unroll 1
            for (int pair = 0; pair < 4; ++pair) {
                if (!((act >> pair) & 1u)) {   // background pair: the precomputed constants (warp-uniform branch)
                    const int z2 = (lane & 7) >> 1;
                    const int j = ((lane & 16) ? 4 : 0) + ((lane & 8) ? 2 : 0) + (lane & 1);
                    const int cls = (pair == 0 ? 0 : (pair == 3 ? 2 : 1)) * 9 + (q == 0 ? 0 : (q == 3 ? 2 : 1)) * 3 +
                                    (z2 == 0 ? 0 : (z2 == 3 ? 2 : 1));
                    const int pos = (pair * 4 + q) * 4 + z2;
                    *reinterpret_cast<float2 *>(out + pos * 16 + 2 * j) =
                        *reinterpret_cast<const float2 *>(sm + SM_BGP + (cls * 16 + 2 * j) * 4);
                    continue;
                }
                uint32_t v[32];
                umma::tmem_ld_x32(tbase + ((uint32_t)(32 * q) << 16) + b * 128 + pair * 32, v);
                umma::tmem_ld_wait();
                float m[16];
#pragma unroll
                for (int c = 0; c < 16; c += 2) {   // (hi.hi + lo.hi columns) + (hi.lo columns) + bias, two channels per packed add
                    const float2 s2 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])),
                                                            make_float2(__uint_as_float(v[16 + c]), __uint_as_float(v[17 + c]))),
                                                 *reinterpret_cast<const float2 *>(b2s + c));
                    m[c] = s2.x; m[c + 1] = s2.y;
                }
                // 2x2x2 max-pool: partners differ in x-slice (lane^16), y (lane^8), z (lane^1).  At each exchange a
                // lane sends the half of its channels the partner keeps and receives the half it keeps itself:
                // 16 -> 8 -> 4 -> 2 channels, 14 shuffles instead of 48; the lane with bits (x,y,z) = (lane>>4&1,
                // lane>>3&1, lane&1) ends with channels 8x + 4y + 2z, +1 of the group's pooled cell.
                float m8[8], m4[4];
                const bool bx = lane & 16, by = lane & 8, bz = lane & 1;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float recv = __shfl_xor_sync(0xffffffffu, bx ? m[c] : m[8 + c], 16);
                    m8[c] = fmaxf(bx ? m[8 + c] : m[c], recv);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float recv = __shfl_xor_sync(0xffffffffu, by ? m8[c] : m8[4 + c], 8);
                    m4[c] = fmaxf(by ? m8[4 + c] : m8[c], recv);
                }
                const float r0 = __shfl_xor_sync(0xffffffffu, bz ? m4[0] : m4[2], 1);
                const float r1 = __shfl_xor_sync(0xffffffffu, bz ? m4[1] : m4[3], 1);
                const float o0 = fmaxf(bz ? m4[2] : m4[0], r0), o1 = fmaxf(bz ? m4[3] : m4[1], r1);
                const int j = (bx ? 4 : 0) + (by ? 2 : 0) + (bz ? 1 : 0);   // channel pair index
                const int pos = (pair * 4 + q) * 4 + ((lane & 7) >> 1);
                *reinterpret_cast<float2 *>(out + pos * 16 + 2 * j) = make_float2(fast_tanh(o0), fast_tanh(o1));
            }
            umma::fence_before_thread_sync();
            umma::mbar_arrive(&tempty[b]);
            if (warp == TC_ISSUER + 1 && lane == 0) stamp(i, 4);
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    if (warp == TC_ISSUER) umma::tmem_dealloc(tbase, 256);
}

// ---- conv1 + conv2, two patches per MMA (M = 128) and dx folded into N -------------------------------------------------
// conv12_tc_kernel above is bound by the tensor core's shared-memory operand fetch (ncu round 2: its wavefronts take 79 % of
// the shared-memory pipe, the tensor pipe is 44 % active): with only 16 output channels every M = 64 x N = 32 MMA fetches
// 2 KB of A for 1 KB of B, 616 KB per patch.  Two changes cut the operand bytes per patch to 280 KB:
//   * TWO patches share every MMA (M = 128): the operand volume is [x 8][y 10 (halo)][patch 2][z 10 (halo)] x 16 B, so
//     the 128 rows (y, patch, z) of tap (dy,dz) in slab x are 16 eight-row groups 160 B apart — one K-major descriptor.
//   * the three dx taps are folded into N: slab x is multiplied ONCE per (dy,dz) tap pair by B = [dx=2 | dx=1 | dx=0] x
//     [W_hi | W_lo] (N = 96) and accumulates into the 32-column blocks of the output slices x-1, x, x+1, which lie side by
//     side in TMEM (slabs 0 and 7 use the N = 64 sub-matrix).  The 9 (dy,dz) taps of the hi and of the lo volume are 4 + 4 K-steps
//     of two taps plus ONE step whose two K chunks are tap 8 of the hi and of the lo volume: 9 MMAs per slab.
// A_hi and A_lo both go through the same B: D[:, hi columns] + D[:, lo columns] = (A_hi + A_lo)(W_hi + W_lo).
// Because one MMA now touches three output blocks of which one may be fresh, the pair's 256 accumulator columns are
// initialised by ONE N = 256 MMA (with the background's contribution, see the kernel's set-up) and everything else accumulates.
// One CTA per SM (all 512 TMEM columns = two pairs in flight): 16 producer warps, 1 issuer warp, 8 epilogue warps.  The
// volumes hold conv1 minus its background value, so a slab is multiplied only when one of the two patches has a listed cell
// in it; an output slice pair whose whole neighbourhood is background in both patches is not even read back from TMEM
// (precomputed constants).
// conv1 of a pair, by all 16 producer warps (the tensor side above needs ~4 k cycles per pair, so conv1 has to stay below):
//   pass 1  thread = (patch, (px,py) column, z quarter): ORs the 4x4 occupancy rows around the column and appends the cells
//           whose 4x4x4 window is non-empty to ONE list for both patches (cell index + coordinates, 4 bytes);
//   pass 2  a warp takes 8 listed cells per round as two independent groups of 4 (eight lanes per cell = the pooled
//           sub-positions).  The 8 lanes first assemble the cell's 64-bit window (each loads two of its 16 rows, three
//           xor-shuffle ORs), then a sub-position's 27 neighbourhood bits are three 9-bit (dy,dz) patterns, one per dx, each
//           indexing a table of precomputed partial weight sums (3 x 512 patterns x 8 channels, built per CTA from the 9 x 8
//           partial-sum table): three table rows added up, max over the eight lanes by channel-halving shuffles, tanh, split.
// Only listed cells cost anything (10 % of the cells on the benchmark's patches): the first version built every cell's window
// in pass 1 (2.4 k warp instructions per patch, now ~0.3 k) and ran the two patches on separate halves of the producers (a
// pair took the slower patch's time).
constexpr int P2_PROD_WARPS = 16;
constexpr int P2_PROD = P2_PROD_WARPS * 32;           // warps 0-15: conv1
constexpr int P2_ISSUER = 16;
constexpr int P2_EPI0 = 17, P2_EPI_WARPS = 8;         // warps 17-24: (lane quarter = warp % 4) x (output slice pairs 0-1 / 2-3)
constexpr int P2_THREADS = (P2_EPI0 + P2_EPI_WARPS) * 32;
constexpr int P2_VOL = 8 * 10 * 2 * 10 * 16;          // 25600 B
constexpr int P2_SM_A = 0;                            // [buf 2][hi, lo][P2_VOL]
constexpr int P2_W_S = 2 * 96 * 16;                   // one K-step of B: [chunk 2][row 96][16 B]
constexpr int P2_SM_W = 4 * P2_VOL;
constexpr int P2_SM_Z = P2_SM_W + 5 * P2_W_S;         // B operand of the accumulator initialisation [chunk 2][row 256][16 B]
constexpr int P2_SM_T3 = P2_SM_Z + 2 * 256 * 16;      // conv1 partial sums [ch half 2][dx 3][512 (dy,dz) patterns][4] f32
constexpr int P2_T3_HALF = 3 * 512 * 16;
constexpr int P2_SM_B12 = P2_SM_T3 + 2 * P2_T3_HALF;
constexpr int P2_SM_BG = P2_SM_B12 + (8 + 16) * 4;
constexpr int P2_SM_PK = P2_SM_BG + 32;               // staged occupancy rows [buf 2][patch 2][PK_BYTES]
constexpr int P2_SM_LST = P2_SM_PK + 4 * PK_BYTES;    // listed cells [buf 2][1024] u32: cell | patch<<16 | px<<20 | py<<24 | pz<<28
constexpr int P2_SM_LCNT = P2_SM_LST + 2 * 1024 * 4;  // [buf 2]
constexpr int P2_SM_XS = P2_SM_LCNT + 16;             // ring of 4 pairs x 2 patches x 8 slice flags
constexpr int P2_SM_BGP = P2_SM_XS + 64;
constexpr int P2_SM_BAR = P2_SM_BGP + 27 * 16 * 4;
constexpr int P2_SM_AI = (P2_SM_BAR + 64 + 127) / 128 * 128;   // A operand of the accumulator initialisation [chunk 2][row 128][16 B]
constexpr int P2_SMEM = P2_SM_AI + 2 * 128 * 16;
static_assert(P2_SM_T3 % 16 == 0 && P2_SM_B12 % 16 == 0 && P2_SM_PK % 16 == 0 && P2_SM_LST % 16 == 0 && P2_SM_XS % 8 == 0 &&
              P2_SM_BGP % 16 == 0 && P2_SM_BAR % 8 == 0 && P2_SM_AI % 128 == 0 && P2_SMEM <= 227 * 1024, "smem layout");

// Row of pattern idx inside a 512-row table.  The eight lanes of a quarter warp (= the eight sub-positions of one cell) read
// their rows with one 16-byte load each, and the patterns of a sparse patch are mostly single bits: 2^k is a multiple of 8
// for k >= 3, so with slot = idx six of the nine one-voxel patterns share a 16-byte bank group (measured: 3.4 wavefronts per
// ideal one, half of all the kernel's shared-memory wavefronts).  The low three bits are therefore XORed with a hash of the
// upper six (a bijection); K = 23, shift 2 was the best of a search over the windows of real patches (tools: 2.05 -> 1.23
// wavefronts per quarter-warp load).
__device__ __forceinline__ unsigned t3_slot(unsigned idx) { return idx ^ ((((idx >> 3) * 23u) >> 2) & 7u); }

// conv1 of U x four listed cells per warp (eight lanes each): window -> three table rows -> max-pool -> tanh -> split fp16.
// The U groups advance in lockstep, stage by stage: one group is a chain of ~150 dependent instructions (shared-memory loads,
// two shuffle butterflies), and with only 16 producer warps on the SM it is their latency, not the issue rate, that bounds
// conv1 — two calls one after the other were not interleaved by the compiler (the shuffles keep their order).
template <int U>
__device__ __forceinline__ void conv1_cells(const unsigned *lst, int k, int n, const unsigned char *pk, const unsigned char *t3,
                                            const float4 &bias_lo, const float4 &bias_hi, unsigned char *a_hi, unsigned char *a_lo,
                                            int sub, float bgsub)
{
    const int ix = sub >> 1, yh = sub & 1;
    const int sx = sub >> 2, sy = (sub >> 1) & 1, sz = sub & 1;
    unsigned ent[U], v[U], wlo[U], whi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) ent[u] = (k + 4 * u < n) ? lst[k + 4 * u] : 0u;
    // the cell's window = rows x = 2px-1 .. 2px+2, y = 2py-1 .. 2py+2 (staged with a halo: row (x,y) at [(x+1)*PITCH + y+1]),
    // bits z = 2pz-1 .. 2pz+2; nibble (ix,iy) at bits 16*ix + 4*iy.  This lane loads rows (ix = sub>>1, iy = 2*(sub&1), +1).
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int pp = (ent[u] >> 16) & 1, px = (ent[u] >> 20) & 7, py = (ent[u] >> 24) & 7;
        v[u] = *reinterpret_cast<const unsigned *>(pk + pp * PK_BYTES + ((2 * px + ix) * PK_PITCH + 2 * py + 2 * yh) * 2);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int pz = ent[u] >> 28;
        const unsigned n01 = ((((v[u] & 0xFFFFu) << 1) >> (2 * pz)) & 0xFu) | (((((v[u] >> 16) << 1) >> (2 * pz)) & 0xFu) << 4);
        const unsigned sh = n01 << (16 * (ix & 1) + 8 * yh);
        wlo[u] = ix < 2 ? sh : 0u;
        whi[u] = ix < 2 ? 0u : sh;
    }
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        unsigned rl[U], rh[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            rl[u] = __shfl_xor_sync(0xffffffffu, wlo[u], d);
            rh[u] = __shfl_xor_sync(0xffffffffu, whi[u], d);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { wlo[u] |= rl[u]; whi[u] |= rh[u]; }
    }
    float2 acc2[U][4];
    float4 w0[U][3], w1[U][3];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned long long w2 = (((unsigned long long)whi[u] << 32) | wlo[u]) >> (16 * sx + 4 * sy + sz);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const unsigned g = (unsigned)(w2 >> (16 * dx)) & 0x777u;                       // (dy,dz) bits at 4*dy + dz
            const unsigned idx = (g & 7u) | ((g >> 1) & 0x38u) | ((g >> 2) & 0x1C0u);
            // the empty pattern's row is all zeros (x + 0 = x exactly) and most patterns of a sparse patch are empty: those lanes
            // skip the loads
            w0[u][dx] = w1[u][dx] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx) {
                const unsigned slot = t3_slot(idx);
                w0[u][dx] = *reinterpret_cast<const float4 *>(t3 + (dx * 512 + slot) * 16);
                w1[u][dx] = *reinterpret_cast<const float4 *>(t3 + P2_T3_HALF + (dx * 512 + slot) * 16);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        acc2[u][0] = make_float2(bias_lo.x, bias_lo.y); acc2[u][1] = make_float2(bias_lo.z, bias_lo.w);
        acc2[u][2] = make_float2(bias_hi.x, bias_hi.y); acc2[u][3] = make_float2(bias_hi.z, bias_hi.w);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            acc2[u][0] = __fadd2_rn(acc2[u][0], make_float2(w0[u][dx].x, w0[u][dx].y));
            acc2[u][1] = __fadd2_rn(acc2[u][1], make_float2(w0[u][dx].z, w0[u][dx].w));
            acc2[u][2] = __fadd2_rn(acc2[u][2], make_float2(w1[u][dx].x, w1[u][dx].y));
            acc2[u][3] = __fadd2_rn(acc2[u][3], make_float2(w1[u][dx].z, w1[u][dx].w));
        }
    }
    // max over the eight sub-position lanes, halving the channel set a lane carries at each exchange: lane `sub` ends with
    // channel 4*sx + 2*sy + sz = sub
    float h4[U][4], h2[U][2], o[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float acc[8] = {acc2[u][0].x, acc2[u][0].y, acc2[u][1].x, acc2[u][1].y, acc2[u][2].x, acc2[u][2].y, acc2[u][3].x, acc2[u][3].y};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float recv = __shfl_xor_sync(0xffffffffu, sx ? acc[c] : acc[4 + c], 4);
            h4[u][c] = fmaxf(sx ? acc[4 + c] : acc[c], recv);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float recv = __shfl_xor_sync(0xffffffffu, sy ? h4[u][c] : h4[u][2 + c], 2);
            h2[u][c] = fmaxf(sy ? h4[u][2 + c] : h4[u][c], recv);
        }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float recv = __shfl_xor_sync(0xffffffffu, sz ? h2[u][0] : h2[u][1], 1);
        o[u] = fmaxf(sz ? h2[u][1] : h2[u][0], recv);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (k + 4 * u < n) {
            const int pi = ent[u] & 0xFFFFu;
            __half h, l;
            umma::split_f16(fast_tanh(o[u]) - bgsub, h, l);   // the volumes hold conv1 minus its background value
            *reinterpret_cast<__half *>(a_hi + pi * 16 + sub * 2) = h;
            *reinterpret_cast<__half *>(a_lo + pi * 16 + sub * 2) = l;
        }
}

__global__ void __launch_bounds__(P2_THREADS, 1) conv12_pair_kernel(const Conv12Args a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float *b1s = reinterpret_cast<float *>(sm + P2_SM_B12), *b2s = b1s + 8;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + P2_SM_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + P2_SM_BAR + 48);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    // ---- one-time setup ----
    for (int i = tid; i < P2_SM_T3 / 16; i += P2_THREADS) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 4 * PK_BYTES / 16; i += P2_THREADS) reinterpret_cast<uint4 *>(sm + P2_SM_PK)[i] = make_uint4(0, 0, 0, 0);
    // T3[half][dx][(dy,dz) pattern][4 ch] = sum over dy of the (dx,dy) partial sums of the pattern's three dz bits
    for (int e = tid; e < 3 * 512 * 8; e += P2_THREADS) {
        const int c = e & 7, idx = (e >> 3) & 511, dx = e >> 12;
        float s = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) s += a.tables[((dx * 3 + dy) * 8 + ((idx >> (3 * dy)) & 7)) * 8 + c];
        *reinterpret_cast<float *>(sm + P2_SM_T3 + (c >> 2) * P2_T3_HALF + (dx * 512 + t3_slot(idx)) * 16 + (c & 3) * 4) = s;
    }
    for (int e = tid; e < 27 * 16; e += P2_THREADS) reinterpret_cast<float *>(sm + P2_SM_BGP)[e] = a.tables[576 + e];
    if (tid < 8) b1s[tid] = a.b1[tid];
    if (tid < 16) b2s[tid] = a.b2[tid];
    if (tid < 2) reinterpret_cast<int *>(sm + P2_SM_LCNT)[tid] = 0;
    for (int i = tid; i < 2 * 128 * 16 / 16; i += P2_THREADS) reinterpret_cast<uint4 *>(sm + P2_SM_AI)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // The operand volumes hold conv1 MINUS its background value tanh(b1): a cell with an empty window is exactly zero, so a
    // slab without a listed cell contributes nothing and its MMAs are skipped, whatever its neighbours look like.  The
    // background's own contribution to every conv2 output, sum over the taps that stay inside the volume of tanh(b1) . W2,
    // only depends on the position's edge class per axis; it is what the accumulators are INITIALISED with, by one K = 16
    // MMA:  A_init[(y,patch,z)][(dy,dz)] = 1 where tap (dy,dz) stays inside in y and z,  B_init[(dy,dz)][(x, hi|lo, ch)] = the
    // (hi | lo part of the) sum over ci and over the dx that stay inside in x of tanh(b1[ci]) W2[dx,dy,dz][ci][ch].
    for (int e = tid; e < 128 * 9; e += P2_THREADS) {
        const int m = e / 9, t = e % 9, y = m >> 4, z = m & 7, dy = t / 3, dz = t % 3;
        const bool in = (unsigned)(y + dy - 1) < 8u && (unsigned)(z + dz - 1) < 8u;
        *reinterpret_cast<__half *>(sm + P2_SM_AI + (t >> 3) * 2048 + (m >> 3) * 128 + (m & 7) * 16 + (t & 7) * 2) =
            __float2half_rn(in ? 1.0f : 0.0f);
    }
    for (int e = tid; e < 8 * 16 * 9; e += P2_THREADS) {
        const int x = e / 144, ch = (e / 9) % 16, t = e % 9;
        float sum = 0.f;
        for (int dx = 0; dx < 3; ++dx) {
            if ((unsigned)(x + dx - 1) >= 8u) continue;
            for (int ci = 0; ci < 8; ++ci) sum = fmaf(tanhf(a.b1[ci]), a.k2[((dx * 9 + t) * 8 + ci) * 16 + ch], sum);
        }
        __half h, l;
        umma::split_f16(sum, h, l);
        unsigned char *w = sm + P2_SM_Z + (t >> 3) * 4096 + (t & 7) * 2;
        const int nh = x * 32 + ch, nl = nh + 16;
        *reinterpret_cast<__half *>(w + (nh >> 3) * 128 + (nh & 7) * 16) = h;
        *reinterpret_cast<__half *>(w + (nl >> 3) * 128 + (nl & 7) * 16) = l;
    }
    // B: K-step s < 4 holds the (dy,dz) taps j = 2s (chunk 0) and 2s+1 (chunk 1); step 4 holds tap 8 in BOTH chunks (its A chunks
    // are tap 8 of the hi and of the lo volume: one MMA serves both parts).  Row n = blk*32 + h*16 + co with blk = 2 - dx (the
    // output slice x + 1 - dx in ascending order), h = 0: W_hi, 1: W_lo.
    for (int e = tid; e < 27 * 8 * 16; e += P2_THREADS) {
        const int t = e / 128, ci = (e / 16) % 8, co = e % 16;
        const int dx = t / 9, j = t % 9, blk = 2 - dx;
        __half h, l;
        umma::split_f16(a.k2[e], h, l);
        const int nh = blk * 32 + co, nl = nh + 16;
        for (int rep = 0; rep < (j < 8 ? 1 : 2); ++rep) {
            const int s = j < 8 ? j >> 1 : 4, chunk = j < 8 ? j & 1 : rep;
            unsigned char *w = sm + P2_SM_W + s * P2_W_S + chunk * (96 * 16) + ci * 2;
            *reinterpret_cast<__half *>(w + (nh >> 3) * 128 + (nh & 7) * 16) = h;
            *reinterpret_cast<__half *>(w + (nl >> 3) * 128 + (nl & 7) * 16) = l;
        }
    }
    if (warp == P2_ISSUER) umma::tmem_alloc(tmem_slot, 512);
    uint64_t *full = mbar, *tfull = mbar + 2, *tempty = mbar + 4;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&full[b], P2_PROD);
            umma::mbar_init(&tfull[b], 1);
            umma::mbar_init(&tempty[b], P2_EPI_WARPS * 32);
        }
        umma::fence_mbar_init();
    }
    umma::fence_proxy_async();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t sA = umma::smem_u32(sm + P2_SM_A), sW = umma::smem_u32(sm + P2_SM_W), sZ = umma::smem_u32(sm + P2_SM_Z);

    // patch pairs: in frame mode with an even K a pair = two consecutive key points of one scale, and consecutive pairs of
    // a CTA cycle through the scales (their conv1 work differs a lot, their conv2 work does not)
    const bool by_scale = a.K3 != 0 && ((a.K3 / 3) & 1) == 0;
    const int n_pairs = (a.P + 1) / 2;
    const int n_my = (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto patch_of = [&](int i, int g) {
        const int v = (int)blockIdx.x + i * (int)gridDim.x;
        int p = 2 * v + g;
        if (by_scale) {
            const int hp = a.K3 / 2, f = v / hp, r = v - f * hp;
            p = f * a.K3 + (r % 3) * (a.K3 / 3) + 2 * (r / 3) + g;
        }
        return p;
    };
    auto stamp = [&](int i, int slot) {
        if (a.timeline && i < 64) a.timeline[((size_t)blockIdx.x * 64 + i) * 16 + slot] = clock64();
    };
    // output slice pairs (pooled x) that need conv2 in at least one patch of the pair `ord & 3`
    auto active_pairs = [&](int ord) -> unsigned {
        if (!a.skip_bg) return 0xFu;
        const unsigned long long *fp = reinterpret_cast<const unsigned long long *>(sm + P2_SM_XS + 16 * ord);
        const unsigned long long f = fp[0] | fp[1];
        unsigned m = 0;
#pragma unroll
        for (int xs = 0; xs < 8; ++xs) m |= (unsigned)((f >> (8 * xs)) & 1ull) << xs;
        const unsigned near = m | (m << 1) | (m >> 1);
        unsigned act = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) act |= ((near >> (2 * p)) & 3u) ? (1u << p) : 0u;
        return act;
    };

    auto dirty_slabs = [&](int ord) -> unsigned {
        if (!a.skip_bg) return 0xFFu;
        const unsigned long long *fp = reinterpret_cast<const unsigned long long *>(sm + P2_SM_XS + 16 * ord);
        const unsigned long long f = fp[0] | fp[1];
        unsigned m = 0;
#pragma unroll
        for (int xs = 0; xs < 8; ++xs) m |= (unsigned)((f >> (8 * xs)) & 1ull) << xs;
        return m;
    };

    if (warp == P2_ISSUER) {
        // ===== MMA issuer =====
        const uint32_t id96 = umma::idesc_f16_f32(128, 96), id64 = umma::idesc_f16_f32(128, 64), id256 = umma::idesc_f16_f32(128, 256);
        const uint64_t a_base = umma::smem_desc(sA, 0, 160), b_base = umma::smem_desc(sW, 96 * 16, 128);
        const uint64_t z_desc = umma::smem_desc(sZ, 256 * 16, 128), ai_desc = umma::smem_desc(umma::smem_u32(sm + P2_SM_AI), 128 * 16, 128);
        for (int j = 0; j < n_my; ++j) {
            const int b = j & 1, k = j >> 1;
            umma::mbar_wait(&full[b], (uint32_t)(k & 1));
            if (k >= 1) umma::mbar_wait(&tempty[b], (uint32_t)((k - 1) & 1));
            umma::fence_after_thread_sync();
            if (lane == 0) stamp(j, 5);
            const unsigned act = active_pairs(j & 3);
            const unsigned need = dirty_slabs(j & 3);   // slabs with a listed cell in either patch: the others are exactly zero
            if (umma::elect_one()) {
                const uint32_t d0 = tbase + b * 256;
                if (act) umma::mma_f16(d0, ai_desc, z_desc, id256, 0u);   // accumulators := the background's contribution
#pragma unroll 1
                for (int x = 0; x < 8; ++x) {
                    if (!((need >> x) & 1u) || (a.dbg & 1)) continue;
                    const uint32_t d = d0 + (x == 0 ? 0 : x - 1) * 32;
                    const uint32_t idesc = (x == 0 || x == 7) ? id64 : id96;
                    const uint64_t bx = b_base + (uint64_t)(x == 0 ? (32 * 16) >> 4 : 0);   // slab 0: rows 32..95 (dx = 1, 0)
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        const uint64_t ax = a_base + (uint64_t)(((2 * b + part) * P2_VOL + x * 3200) >> 4);
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            // taps (2s, 2s+1): tap j = dy*3+dz sits dy*320 + dz*16 bytes into the slab
                            const int j0 = 2 * s, j1 = j0 + 1;
                            const int o0 = (j0 / 3) * 320 + (j0 % 3) * 16, o1 = (j1 / 3) * 320 + (j1 % 3) * 16;
                            const uint64_t da = ax + (uint64_t)(o0 >> 4) + ((uint64_t)((o1 - o0) >> 4) << 16);
                            const uint64_t db = bx + (uint64_t)((s * P2_W_S) >> 4);
                            umma::mma_f16(d, da, db, idesc, 1u);
                        }
                    }
                    {   // tap 8 of the hi volume (chunk 0) and of the lo volume (chunk 1, LBO = the distance between the volumes)
                        const uint64_t da = a_base + (uint64_t)(((2 * b) * P2_VOL + x * 3200 + 2 * 320 + 2 * 16) >> 4) + ((uint64_t)(P2_VOL >> 4) << 16);
                        umma::mma_f16(d, da, bx + (uint64_t)((4 * P2_W_S) >> 4), idesc, 1u);
                    }
                }
                umma::commit(&tfull[b]);
            }
            __syncwarp();
            if (lane == 0) stamp(j, 6);
        }
    } else if (warp < P2_ISSUER) {
        // ===== producers (16 warps): conv1 of both patches of pair i into operand buffer i&1 =====
        const int g = tid >> 8, t = tid & 255;          // pass 1: this thread's patch of the pair, its thread within the patch
        unsigned *lst_all = reinterpret_cast<unsigned *>(sm + P2_SM_LST);
        int *lcnt = reinterpret_cast<int *>(sm + P2_SM_LCNT);
        unsigned pk_next = 0u;
        int n_dirty0 = 0, n_dirty1 = 0;
        const float4 bias_lo = *reinterpret_cast<const float4 *>(b1s), bias_hi = *reinterpret_cast<const float4 *>(b1s + 4);
        const float bgsub = tanhf(b1s[lane & 7]);       // background value of the channel this lane finishes in pass 2
        auto fetch = [&](int i) {
            if (i < n_my) {
                int p = patch_of(i, g);
                if (p >= a.P) p = a.P - 1;              // odd patch count: the last pair's second half repeats the first
                if (t < 128) pk_next = __ldg(a.packed + (size_t)p * 128 + t);
            }
        };
        fetch(0);
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            if (tid == 0) stamp(i, 0);
            if (i >= 2) {   // the MMAs of pair i-2 must have finished reading operand buffer b
                umma::mbar_wait(&tfull[b], (uint32_t)(((i - 2) >> 1) & 1));
                umma::fence_after_thread_sync();
            }
            unsigned char *a_hi = sm + P2_SM_A + (2 * b) * P2_VOL, *a_lo = a_hi + P2_VOL;
            unsigned *lst = lst_all + b * 1024;
            // the cells the previous pair of this buffer wrote go back to zero (= background)
            const int nd = b ? n_dirty1 : n_dirty0;
            for (int k = tid; k < nd; k += P2_PROD) {
                const int pi = lst[k] & 0xFFFFu;
                *reinterpret_cast<uint4 *>(a_hi + pi * 16) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(a_lo + pi * 16) = make_uint4(0, 0, 0, 0);
            }
            // (staged rows are double-buffered like the operands: other warps may still be in pass 2 of the previous pair)
            unsigned short *rows = reinterpret_cast<unsigned short *>(sm + P2_SM_PK + (2 * b + g) * PK_BYTES);
            if (t < 128) {  // word t = occupancy rows (x, y) and (x, y+1), x = t/8, y = 2*(t%8)
                unsigned short *dst = rows + ((t >> 3) + 1) * PK_PITCH + 2 * (t & 7) + 1;
                dst[0] = (unsigned short)(pk_next & 0xFFFFu);
                dst[1] = (unsigned short)(pk_next >> 16);
            }
            fetch(i + 1);
            asm volatile("bar.sync 1, 512;" ::: "memory");   // rows staged (lcnt[b] was reset during the previous pair's pass 2)
            long long *tl = (a.timeline && i < 64) ? a.timeline + ((size_t)blockIdx.x * 64 + i) * 16 : nullptr;
            if (tl && tid == 0) tl[2] = clock64();
            // ---- pass 1: list the cells with a non-empty window ----
            unsigned found[2];
            {
                const int col = t >> 2, zq = t & 3, px = col >> 3, py = col & 7;
                const unsigned *rp = reinterpret_cast<const unsigned *>(rows + (2 * px) * PK_PITCH + 2 * py);
                unsigned any = 0;
#pragma unroll
                for (int ix = 0; ix < 4; ++ix) any |= rp[ix * (PK_PITCH / 2)] | rp[ix * (PK_PITCH / 2) + 1];
                any = ((any | (any >> 16)) & 0xFFFFu) << 1;   // bit z+1 <-> some row of the column has voxel z
                unsigned cells = 0;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int pz = 2 * zq + half;
                    const bool nz = ((any >> (2 * pz)) & 0xFu) != 0u;   // a voxel at z = 2pz-1 .. 2pz+2
                    found[half] = __ballot_sync(0xffffffffu, nz);
                    cells |= found[half];
                }
                if (lane == 0) sm[P2_SM_XS + 16 * (i & 3) + 8 * g + (t >> 5)] = cells != 0u;   // the warp = x-slice px of patch g
                const int cnt = __popc(found[0]) + __popc(found[1]);
                if (cnt) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(lcnt + b, cnt);
                    base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if ((found[half] >> lane) & 1u) {
                            const int pz = 2 * zq + half;
                            const int k = base + __popc(found[half] & ((1u << lane) - 1u)) + (half ? __popc(found[0]) : 0);
                            lst[k] = (unsigned)(((px * 10 + py + 1) * 2 + g) * 10 + pz + 1) | ((unsigned)g << 16) | ((unsigned)px << 20) |
                                     ((unsigned)py << 24) | ((unsigned)pz << 28);
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (tl && tid == 0) tl[7] = clock64();
            // ---- pass 2: eight listed cells per warp and round, as two independent groups of four ----
            const int n = (a.dbg & 2) ? 0 : lcnt[b];
            if (tid == 0) lcnt[b ^ 1] = 0;   // the next pair's counter: last read before this pair's first barrier, next used after its own
            if (b) n_dirty1 = n; else n_dirty0 = n;
            const int sub = lane & 7;
            for (int k0 = warp * 8 + (lane >> 3); k0 - (lane >> 3) < n; k0 += P2_PROD_WARPS * 8) {
                conv1_cells<2>(lst, k0, n, sm + P2_SM_PK + 2 * b * PK_BYTES, sm + P2_SM_T3, bias_lo, bias_hi, a_hi, a_lo, sub, bgsub);
            }
            if (tl && tid == 0) tl[8] = clock64();
            umma::fence_proxy_async();
            umma::mbar_arrive(&full[b]);
            if (tid == 0) stamp(i, 1);
        }
    } else {
        // ===== epilogue (8 warps): TMEM lane = (y, patch, z); warp -> lane quarter q (pooled y = q) and two of the four
        // output slice pairs =====
        const int q = warp & 3, xh = (warp - P2_EPI0) >> 2;
        const int yb = (lane >> 4) & 1, pp = (lane >> 3) & 1, zb = lane & 1, z2 = (lane & 7) >> 1;
        const int jq = yb * 2 + zb;                      // the lane ends with channels 4*jq .. 4*jq+3 of its pooled cell
        const float4 bias4 = *reinterpret_cast<const float4 *>(b2s + 4 * jq);
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            umma::mbar_wait(&tfull[b], (uint32_t)((i >> 1) & 1));
            umma::fence_after_thread_sync();
            if (warp == P2_EPI0 && lane == 0) stamp(i, 3);
            const int p = patch_of(i, pp);
            const bool valid = p < a.P;
            float *out = a.act2 + (size_t)(valid ? p : 0) * 1024;
            const unsigned act = active_pairs(i & 3);
#pragma unroll 1
            for (int px = 2 * xh; px < 2 * xh + 2; ++px) {
                const int pos = (px * 4 + q) * 4 + z2;
                if (!((act >> px) & 1u)) {
                    const int cls = (px == 0 ? 0 : (px == 3 ? 2 : 1)) * 9 + (q == 0 ? 0 : (q == 3 ? 2 : 1)) * 3 +
                                    (z2 == 0 ? 0 : (z2 == 3 ? 2 : 1));
                    if (valid)
                        *reinterpret_cast<float4 *>(out + pos * 16 + 4 * jq) =
                            *reinterpret_cast<const float4 *>(sm + P2_SM_BGP + (cls * 16 + 4 * jq) * 4);
                    continue;
                }
                const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16) + b * 256 + px * 64;
                uint32_t v[32];
                float m[16];
                umma::tmem_ld_x32(trow, v);              // slice 2px: 16 hi + 16 lo columns
                umma::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) m[c] = __uint_as_float(v[c]) + __uint_as_float(v[16 + c]);
                umma::tmem_ld_x32(trow + 32, v);         // slice 2px+1
                umma::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) m[c] = fmaxf(m[c], __uint_as_float(v[c]) + __uint_as_float(v[16 + c]));
                // y partner = lane ^ 16, z partner = lane ^ 1; every exchange halves the channels a lane carries
                float m8[8], m4[4];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float recv = __shfl_xor_sync(0xffffffffu, yb ? m[c] : m[8 + c], 16);
                    m8[c] = fmaxf(yb ? m[8 + c] : m[c], recv);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float recv = __shfl_xor_sync(0xffffffffu, zb ? m8[c] : m8[4 + c], 1);
                    m4[c] = fmaxf(zb ? m8[4 + c] : m8[c], recv);
                }
                if (valid)   // the bias is common to the pooled cells: added after the max
                    *reinterpret_cast<float4 *>(out + pos * 16 + 4 * jq) =
                        make_float4(fast_tanh(m4[0] + bias4.x), fast_tanh(m4[1] + bias4.y), fast_tanh(m4[2] + bias4.z),
                                    fast_tanh(m4[3] + bias4.w));
            }
            umma::fence_before_thread_sync();
            umma::mbar_arrive(&tempty[b]);
            if (warp == P2_EPI0 && lane == 0) stamp(i, 4);
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    if (warp == P2_ISSUER) umma::tmem_dealloc(tbase, 512);
}

// ---- conv3 (tcgen05) ---------------------------------------------------------------------------
// Per patch an IMPLICIT GEMM  D[64 positions, 32 ch] = sum over 27 taps of A_tap[64,16] W_tap[16,32]:
// the 4^3 x 16ch input is stored as 9 (dy,dz)-shifted COMPACT copies of a [6(x halo)][4][4] volume per
// channel half, so that for every tap the 64 output rows are one contiguous run of 16-byte rows
// (canonical K-major layout, SBO = 128 B) starting dx*256 B into copy (dy,dz); the two channel halves
// are the two K chunks (LBO = 9*1536 B).  Split fp16: A_hi x [W_hi|W_lo] (N=64) + A_lo x W_hi (N=32),
// fp32 accumulators in TMEM (2 slots x 64 columns), epilogue adds the halves, bias, tanh, and writes
// act3 as split fp16 (the dense1 operands).  8 worker warps + 1 MMA-issuer warp, mbarrier pipeline.
constexpr int C3_ISSUER = 12;                        // warps 0-3 produce operands, 4-11 drain accumulators, 12 issues the MMAs
constexpr int C3_THREADS = (C3_ISSUER + 1) * 32;
constexpr int C3_COPY = 6 * 16 * 16;                 // 1536 B: [xi 6][y 4][z 4] x 16 B
constexpr int C3_HALF = 9 * C3_COPY;                 // 13824 B: 9 (dy,dz) copies of one channel half
constexpr int C3_PART = 2 * C3_HALF;                 // 27648 B: both halves
constexpr int C3_BUF = 2 * C3_PART;                  // 55296 B: hi + lo
// NBUF operand buffers: 2 = one CTA per SM, produce(i+1) overlaps MMA(i) inside the CTA; 1 = two CTAs per SM
// (110.8 KB each) that overlap each other — twice the worker warps to hide the producer/epilogue latencies
constexpr int SM3_C = 0;
__host__ __device__ constexpr int sm3_w(int nbuf) { return nbuf * C3_BUF; }                 // W3 [kc 54][n 64][16 B]
__host__ __device__ constexpr int sm3_b3(int nbuf) { return sm3_w(nbuf) + 54 * 1024; }
__host__ __device__ constexpr int sm3_bar(int nbuf) { return sm3_b3(nbuf) + 128; }
__host__ __device__ constexpr int c3_smem(int nbuf) { return sm3_bar(nbuf) + 64; }

struct Conv3Args {
    const float *act2;     // [P,64,16]
    const float *k3, *b3;  // (27,16,32), (32)
    __half *act3_hi, *act3_lo;  // split fp16 dense1 operands, tile-major: [ceil(P/256)][256 chunks][256 patches][8]
    int P;
    int dbg;                    // measurement only (CAELO_CONV3_DBG, conv3_oct_kernel): bit 0 = producers write nothing, bit 1 = the
                                // epilogue only drains, bit 2 = no MMAs (results are wrong)
};

template <int NBUF>
__global__ void __launch_bounds__(C3_THREADS, 3 - NBUF) conv3_tc_kernel(const Conv3Args a)
{
    constexpr int SM3_W = sm3_w(NBUF), SM3_B3 = sm3_b3(NBUF), SM3_BAR = sm3_bar(NBUF);
    extern __shared__ __align__(128) unsigned char sm[];
    float *b3s = reinterpret_cast<float *>(sm + SM3_B3);
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + SM3_BAR), *tfull = full + 2, *tempty = full + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + SM3_BAR + 48);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    for (int i = tid; i < SM3_W / 16; i += C3_THREADS) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    // B operand: row n (0..31 = W_hi, 32..63 = W_lo of out-channel n%32), chunk kc = tap*2 + ci/8
    for (int e = tid; e < 27 * 16 * 32; e += C3_THREADS) {
        int t = e / 512, ci = (e / 32) % 16, co = e % 32;
        __half h, l;
        umma::split_f16(a.k3[e], h, l);
        unsigned char *w = sm + SM3_W + (t * 2 + ci / 8) * 1024 + (ci % 8) * 2;
        *reinterpret_cast<__half *>(w + (co / 8) * 128 + (co % 8) * 16) = h;
        *reinterpret_cast<__half *>(w + ((32 + co) / 8) * 128 + (co % 8) * 16) = l;
    }
    if (tid < 32) b3s[tid] = a.b3[tid];
    if (warp == C3_ISSUER) umma::tmem_alloc(tmem_slot, 128);
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&full[b], 128);     // the four producer warps
            umma::mbar_init(&tfull[b], 1);
            umma::mbar_init(&tempty[b], 256);   // the eight epilogue warps
        }
        umma::fence_mbar_init();
    }
    umma::fence_proxy_async();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t sC = umma::smem_u32(sm + SM3_C), sW = umma::smem_u32(sm + SM3_W);
    const int n_my = (a.P - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto patch_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };

    if (warp == C3_ISSUER) {
        const uint32_t idesc64 = umma::idesc_f16_f32(64, 64), idesc32 = umma::idesc_f16_f32(64, 32);
        for (int j = 0; j < n_my; ++j) {
            const int b = j & 1, k = j >> 1;
            umma::mbar_wait(&full[b], (uint32_t)(k & 1));
            if (k >= 1) umma::mbar_wait(&tempty[b], (uint32_t)((k - 1) & 1));
            umma::fence_after_thread_sync();
            if (umma::elect_one()) {
                const uint32_t d = tbase + b * 64;
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    const uint32_t cb = sC + (NBUF == 2 ? b : 0) * C3_BUF + part * C3_PART;
#pragma unroll
                    for (int t = 0; t < 27; ++t) {
                        const int dx = t / 9, dy = (t / 3) % 3, dz = t % 3;
                        uint64_t da = umma::smem_desc(cb + (dy * 3 + dz) * C3_COPY + dx * 256, C3_HALF, 128);
                        uint64_t db = umma::smem_desc(sW + t * 2048, 1024, 128);
                        umma::mma_f16(d, da, db, part ? idesc32 : idesc64, (part | t) ? 1u : 0u);
                    }
                }
                umma::commit(&tfull[b]);
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // ===== producers (4 warps): scatter one patch's activations (split fp16) into the 9 shifted copies of
        // buffer b.  Thread = (position, channel half): it writes the hi AND the lo part.  The global loads of the
        // NEXT patch are issued before this patch's stores, so their latency hides behind the conversion.
        const int pos = tid >> 1, half = tid & 1;
        float4 nv0 = make_float4(0.f, 0.f, 0.f, 0.f), nv1 = nv0;
        auto fetch = [&](int i) {
            if (i >= n_my) return;
            const float4 *src = reinterpret_cast<const float4 *>(a.act2 + (size_t)patch_of(i) * 1024 + pos * 16 + half * 8);
            nv0 = __ldg(src);
            nv1 = __ldg(src + 1);
        };
        fetch(0);
        const int x = pos >> 4, y = (pos >> 2) & 3, z = pos & 3;
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            const float4 v0 = nv0, v1 = nv1;
            fetch(i + 1);
            if (i >= 2) {   // MMA(i-2) must have finished reading buffer b
                umma::mbar_wait(&tfull[b], (uint32_t)(((i - 2) >> 1) & 1));
                umma::fence_after_thread_sync();
            }
            const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            __half2 hv[4], lv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) umma::split_f16x2(f[2 * c], f[2 * c + 1], hv[c], lv[c]);
            const uint4 vh = *reinterpret_cast<uint4 *>(hv), vl = *reinterpret_cast<uint4 *>(lv);
            unsigned char *base = sm + SM3_C + (NBUF == 2 ? b : 0) * C3_BUF + half * C3_HALF;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
                    const int ys = y + 1 - dy, zs = z + 1 - dz;
                    if ((unsigned)ys < 4u && (unsigned)zs < 4u) {
                        unsigned char *p = base + (dy * 3 + dz) * C3_COPY + (((x + 1) * 4 + ys) * 4 + zs) * 16;
                        *reinterpret_cast<uint4 *>(p) = vh;
                        *reinterpret_cast<uint4 *>(p + C3_PART) = vl;
                    }
                }
            umma::fence_proxy_async();
            umma::mbar_arrive(&full[b]);
        }
    } else {
        // ===== epilogue (8 warps = TMEM lane quarter x channel half): hi/lo halves added, bias, tanh, act3 as split fp16 =====
        const int q = warp & 3, hc = (warp - 4) >> 2;
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            umma::mbar_wait(&tfull[b], (uint32_t)((i >> 1) & 1));
            umma::fence_after_thread_sync();
            uint32_t v0[16], v1[16];
            const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16) + b * 64 + hc * 16;
            umma::tmem_ld_x16(trow, v0);        // columns 0..31: W_hi (A_hi + A_lo products)
            umma::tmem_ld_x16(trow + 32, v1);   // columns 32..63: W_lo
            umma::tmem_ld_wait();
            umma::fence_before_thread_sync();
            umma::mbar_arrive(&tempty[b]);
            if (lane < 16) {  // M=64 accumulator: rows 16q..16q+15 live in lanes 0..15 of quarter q
                const int pos = 16 * q + lane;
                // act3 is stored TILE-MAJOR for dense1: [tile of 256 patches][chunk of 8 k (256 of them)][patch][8 halves],
                // so that a K stage of a dense1 CTA is one contiguous block (a single bulk copy, no 16-byte gather)
                const int p = patch_of(i);
                const size_t row8 = ((size_t)(p >> 8) * 256 * 256 + (size_t)(p & 255)) * 8;   // halves
                const int c0 = pos * 4 + hc * 2;                                               // chunk of channel hc*16 at this position
#pragma unroll
                for (int g = 0; g < 2; ++g) {   // 8 channels at a time
                    __half2 hh[4], ll[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int ch = 8 * g + 2 * c;
                        const float o0 = fast_tanh((__uint_as_float(v0[ch]) + __uint_as_float(v1[ch])) + b3s[hc * 16 + ch]);
                        const float o1 = fast_tanh((__uint_as_float(v0[ch + 1]) + __uint_as_float(v1[ch + 1])) + b3s[hc * 16 + ch + 1]);
                        umma::split_f16x2(o0, o1, hh[c], ll[c]);
                    }
                    const size_t o = row8 + (size_t)(c0 + g) * 256 * 8;
                    *reinterpret_cast<uint4 *>(a.act3_hi + o) = *reinterpret_cast<uint4 *>(hh);
                    *reinterpret_cast<uint4 *>(a.act3_lo + o) = *reinterpret_cast<uint4 *>(ll);
                }
            }
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    if (warp == C3_ISSUER) umma::tmem_dealloc(tbase, 128);
}

// ---- conv3, eight patches per MMA (M = 128) and dx folded into N ---------------------------------------------------------
// conv3_tc_kernel is bound by the tensor core's shared-memory operand fetch (80 % of the shared-memory pipe at 189 KB per
// patch: an M = 64 MMA fetches 2 KB of A for 2 or 1 KB of B).  The same two moves as in conv12_pair_kernel cut that to 90 KB:
//   * rows = (patch 8, y 4, z 4) of ONE x slab: M = 128, all TMEM lanes used.  For tap (dy,dz) the slab is stored as the
//     (dy,dz)-shifted compact copy [row 128][16 B] per channel half (zero where the shifted position leaves the 4 x 4 plane),
//     so every tap's A operand is one contiguous canonical K-major block;
//   * the three dx taps are folded into N: slab x is multiplied once per (dy,dz) by B = [dx=2 | dx=1 | dx=0] x [W_hi | W_lo]
//     (N = 192) and accumulates into the 64-column blocks of the output slices x-1, x, x+1 (slabs 0 and 3: the N = 128
//     sub-matrix).  A_hi and A_lo go through the same B: D[:, hi] + D[:, lo] = (A_hi + A_lo)(W_hi + W_lo).
// 72 MMAs of 10 KB per eight patches.  One slab of one part (hi or lo) with its nine copies is a 36 KB STAGE; four stages form
// a ring that the producers refill (hi part of slabs 0-3, then the lo part) as the MMAs release them.  The group's 256
// accumulator columns are cleared by one N = 256 MMA against a zero B operand; two groups' accumulators fit the 512 TMEM
// columns, so the epilogue of group g runs under the MMAs of group g+1.
constexpr int C8_PROD_WARPS = 8;                       // warps 0-7: thread = (patch 8, position-in-slab 16, channel half 2)
constexpr int C8_ISSUER = 8;
constexpr int C8_EPI0 = 9, C8_EPI_WARPS = 16;          // warps 9-24: (lane quarter = warp % 4) x (output slice)
constexpr int C8_THREADS = (C8_EPI0 + C8_EPI_WARPS) * 32;
constexpr int C8_COPY = 2 * 128 * 16;                  // one (dy,dz) copy of a slab: [channel half 2][row 128][16 B]
constexpr int C8_STAGE = 9 * C8_COPY;                  // 36864 B
constexpr int C8_SM_W = 4 * C8_STAGE;                  // B: [tap 9][chunk 2][row 192][16 B]
constexpr int C8_W_TAP = 2 * 192 * 16;
constexpr int C8_SM_Z = C8_SM_W + 9 * C8_W_TAP;        // zero B operand [chunk 2][row 256][16 B]
constexpr int C8_SM_B3 = C8_SM_Z + 2 * 256 * 16;
constexpr int C8_SM_BAR = C8_SM_B3 + 128;              // full[4] empty[4] tfull[2] tempty[2] + tmem slot
constexpr int C8_SM_STG = C8_SM_BAR + 12 * 8 + 16;     // epilogue store staging [slice team 4][hi, lo][position 16][patch 8][16 B]
constexpr int C8_SMEM = C8_SM_STG + 4 * 4096;
static_assert(C8_SMEM <= 227 * 1024, "conv3_oct_kernel shared memory");

__global__ void __launch_bounds__(C8_THREADS, 1) conv3_oct_kernel(const Conv3Args a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float *b3s = reinterpret_cast<float *>(sm + C8_SM_B3);
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + C8_SM_BAR), *empty = full + 4, *tfull = full + 8, *tempty = full + 10;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + C8_SM_BAR + 12 * 8);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    // stages and the zero operand start as zeros (rows a copy never writes are its zero padding); B = the weights
    for (int i = tid; i < C8_SM_W / 16; i += C8_THREADS) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 2 * 256 * 16 / 16; i += C8_THREADS) reinterpret_cast<uint4 *>(sm + C8_SM_Z)[i] = make_uint4(0, 0, 0, 0);
    // B row n = blk*64 + h*32 + co with blk = 2 - dx (output slice x + 1 - dx ascending), h = 0: W_hi, 1: W_lo; chunk = ci/8
    for (int e = tid; e < 27 * 16 * 32; e += C8_THREADS) {
        const int t = e / 512, ci = (e / 32) % 16, co = e % 32;
        const int dx = t / 9, c = t % 9, blk = 2 - dx;
        __half h, l;
        umma::split_f16(a.k3[e], h, l);
        unsigned char *w = sm + C8_SM_W + c * C8_W_TAP + (ci >> 3) * (192 * 16) + (ci & 7) * 2;
        const int nh = blk * 64 + co, nl = nh + 32;
        *reinterpret_cast<__half *>(w + (nh >> 3) * 128 + (nh & 7) * 16) = h;
        *reinterpret_cast<__half *>(w + (nl >> 3) * 128 + (nl & 7) * 16) = l;
    }
    if (tid < 32) b3s[tid] = a.b3[tid];
    if (warp == C8_ISSUER) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) {
            umma::mbar_init(&full[s], C8_PROD_WARPS * 32);
            umma::mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&tfull[b], 1);
            umma::mbar_init(&tempty[b], C8_EPI_WARPS * 32);
        }
        umma::fence_mbar_init();
    }
    umma::fence_proxy_async();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t sS = umma::smem_u32(sm), sW = umma::smem_u32(sm + C8_SM_W), sZ = umma::smem_u32(sm + C8_SM_Z);
    const int n_groups = (a.P + 7) / 8;
    const int n_my = (n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // groups of this CTA
    auto group_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };

    if (warp == C8_ISSUER) {
        const uint32_t id192 = umma::idesc_f16_f32(128, 192), id128 = umma::idesc_f16_f32(128, 128), id256 = umma::idesc_f16_f32(128, 256);
        const uint64_t a_base = umma::smem_desc(sS, 128 * 16, 128), b_base = umma::smem_desc(sW, 192 * 16, 128);
        const uint64_t z_desc = umma::smem_desc(sZ, 256 * 16, 128);
        for (int j = 0; j < n_my; ++j) {
            const int b = j & 1;
            if (j >= 2) umma::mbar_wait(&tempty[b], (uint32_t)(((j >> 1) - 1) & 1));
            const uint32_t d0 = tbase + b * 256;
#pragma unroll
            for (int f = 0; f < 8; ++f) {            // fill f = (part = f / 4, slab x = f % 4) of this group
                const int x = f & 3;
                umma::mbar_wait(&full[x], (uint32_t)((2 * j + (f >> 2)) & 1));
                umma::fence_after_thread_sync();
                if (umma::elect_one()) {
                    if (f == 0) umma::mma_f16(d0, a_base, z_desc, id256, 0u);   // clear the group's accumulators
                    const uint32_t d = d0 + (x == 0 ? 0 : x - 1) * 64;
                    const uint32_t idesc = (x == 0 || x == 3) ? id128 : id192;
                    const uint64_t bx = b_base + (uint64_t)(x == 0 ? (64 * 16) >> 4 : 0);   // slab 0: rows 64..191 (dx = 1, 0)
                    const uint64_t ax = a_base + (uint64_t)((x * C8_STAGE) >> 4);
                    if (!(a.dbg & 4))
#pragma unroll
                    for (int c = 0; c < 9; ++c)
                        umma::mma_f16(d, ax + (uint64_t)((c * C8_COPY) >> 4), bx + (uint64_t)((c * C8_W_TAP) >> 4), idesc, 1u);
                    umma::commit(&empty[x]);
                    if (f == 7) umma::commit(&tfull[b]);
                }
                __syncwarp();
            }
        }
    } else if (warp < C8_ISSUER) {
        // ===== producers: thread = (patch p of the group, position yz of a slab, channel half).  Per group it loads its eight
        // floats of every slab (the next group's loads are issued before this group's stores), splits them once, and writes
        // the hi part of slabs 0-3 and then the lo part into the nine shifted copies of the slab's stage, each time after the
        // MMAs that read the stage's previous content have completed =====
        const int p8 = tid >> 5, yz = tid & 15, half = (tid >> 4) & 1;   // a quarter warp = 8 consecutive rows of one half: no bank conflicts
        const int y = yz >> 2, z = yz & 3;
        float4 nv[4][2];
        auto fetch = [&](int i) {
            if (i >= n_my) return;
            int p = group_of(i) * 8 + p8;
            if (p >= a.P) p = a.P - 1;                 // a short last group repeats its last patch (never stored)
            const float4 *src = reinterpret_cast<const float4 *>(a.act2 + (size_t)p * 1024 + yz * 16 + half * 8);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                nv[x][0] = __ldg(src + x * 64);        // slab x: positions x*16 + yz, 16 floats each
                nv[x][1] = __ldg(src + x * 64 + 1);
            }
        };
        fetch(0);
        // destination rows of this thread's value in the nine copies: copy (dy,dz) holds input (y,z) at row (y+1-dy, z+1-dz)
        for (int i = 0; i < n_my; ++i) {
            uint4 vh[4], vl[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const float f[8] = {nv[x][0].x, nv[x][0].y, nv[x][0].z, nv[x][0].w, nv[x][1].x, nv[x][1].y, nv[x][1].z, nv[x][1].w};
                __half2 hv[4], lv[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) umma::split_f16x2(f[2 * c], f[2 * c + 1], hv[c], lv[c]);
                vh[x] = *reinterpret_cast<uint4 *>(hv);
                vl[x] = *reinterpret_cast<uint4 *>(lv);
            }
            fetch(i + 1);
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                const int x = f & 3, part = f >> 2;
                const int fill = 2 * i + part;           // how many times stage x has been filled before
                if (fill >= 1) {
                    umma::mbar_wait(&empty[x], (uint32_t)((fill - 1) & 1));
                    umma::fence_after_thread_sync();
                }
                const uint4 v = part ? vl[x] : vh[x];
                unsigned char *st = sm + x * C8_STAGE + half * (128 * 16);
                if (!(a.dbg & 1))
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz) {
                        const int ys = y + 1 - dy, zs = z + 1 - dz;
                        if ((unsigned)ys < 4u && (unsigned)zs < 4u)
                            *reinterpret_cast<uint4 *>(st + (dy * 3 + dz) * C8_COPY + (p8 * 16 + ys * 4 + zs) * 16) = v;
                    }
                umma::fence_proxy_async();
                umma::mbar_arrive(&full[x]);
            }
        }
    } else {
        // ===== epilogue (16 warps): TMEM lane = (patch, y, z); warp -> lane quarter q (patches 2q, 2q+1) and ONE output slice.
        // act3 is tile-major for dense1 ([tile of 256 patches][chunk of 8 k][patch][16 B]): for one chunk the group's eight
        // patches are one 128-byte line, but a warp holds 2 patches x 16 positions, i.e. 32 lines.  Written straight from the
        // registers that is 32 partial sectors per store instruction, and with the operands cheap those stores bounded the
        // kernel (0.34 ms with the epilogue only draining, 0.53 ms with it, whether 8 or 16 warps).  So the four warps of a
        // slice exchange their 16-byte pieces through shared memory (swizzled by position: conflict-free both ways) and every
        // quarter warp writes one full line. =====
        const int q = warp & 3, x = (warp - C8_EPI0) >> 2;
        const int pl = 2 * q + (lane >> 4), yz = lane & 15;
        unsigned char *stg = sm + C8_SM_STG + x * 4096;
        unsigned char *my_slot = stg + yz * 128 + ((pl ^ (yz & 7)) << 4);
        const int tt = q * 32 + lane;                          // thread of the slice team: writes line tt/8, patch tt%8
        const int wl = tt >> 3, wj = tt & 7;
        const unsigned char *rd_slot = stg + wl * 128 + ((wj ^ (wl & 7)) << 4);
        const int barid = 1 + x;
        for (int i = 0; i < n_my; ++i) {
            const int b = i & 1;
            umma::mbar_wait(&tfull[b], (uint32_t)((i >> 1) & 1));
            umma::fence_after_thread_sync();
            const int pw = group_of(i) * 8 + wj;               // the patch this thread writes out
            const size_t row8 = ((size_t)(pw >> 8) * 256 * 256 + (size_t)(pw & 255)) * 8;   // halves (tile-major act3)
            const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16) + b * 256 + x * 64;
#pragma unroll
            for (int g = 0; g < 4; ++g) {                 // 8 channels at a time: chunk pos*4 + g of the dense1 K axis
                uint32_t v0[8], v1[8];
                umma::tmem_ld_x8(trow + g * 8, v0);        // W_hi columns
                umma::tmem_ld_x8(trow + 32 + g * 8, v1);   // W_lo columns
                umma::tmem_ld_wait();
                __half2 hh[4], ll[4];
                if (!(a.dbg & 2)) {
                    const float4 c0 = *reinterpret_cast<const float4 *>(b3s + 8 * g), c1 = *reinterpret_cast<const float4 *>(b3s + 8 * g + 4);
                    const float bb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float o0 = fast_tanh((__uint_as_float(v0[2 * c]) + __uint_as_float(v1[2 * c])) + bb[2 * c]);
                        const float o1 = fast_tanh((__uint_as_float(v0[2 * c + 1]) + __uint_as_float(v1[2 * c + 1])) + bb[2 * c + 1]);
                        umma::split_f16x2(o0, o1, hh[c], ll[c]);
                    }
                    *reinterpret_cast<uint4 *>(my_slot) = *reinterpret_cast<uint4 *>(hh);
                    *reinterpret_cast<uint4 *>(my_slot + 2048) = *reinterpret_cast<uint4 *>(ll);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
                if (pw < a.P && !(a.dbg & 2)) {
                    const size_t o = row8 + (size_t)((x * 16 + wl) * 4 + g) * 256 * 8;
                    *reinterpret_cast<uint4 *>(a.act3_hi + o) = *reinterpret_cast<const uint4 *>(rd_slot);
                    *reinterpret_cast<uint4 *>(a.act3_lo + o) = *reinterpret_cast<const uint4 *>(rd_slot + 2048);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
            }
            umma::fence_before_thread_sync();
            umma::mbar_arrive(&tempty[b]);
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    if (warp == C8_ISSUER) umma::tmem_dealloc(tbase, 512);
}

// ---- dense1 (tcgen05) + tanh + dense2 + tanh ---------------------------------------------------
// One CTA = 256 patches (two M = 128 tiles) x 208 outputs (N = 200 padded) x K = 2048, split-fp16 operands:
//   D_t += A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T   (one fp32 accumulator per tile in TMEM, 2 x 208 columns),
//   h = tanh(D_t + b).  The kernel is bound by the L2 -> smem operand stream (the 1.7 MB of weights are
//   re-streamed per CTA), hence the tall CTA tile: weights are fetched once per 256 patches.
// Operands stream to smem by 1-D bulk copies (cp.async.bulk + mbarrier complete_tx).  Both are stored TILE-MAJOR in
// global memory — act3 by conv3's epilogue as [tile of 256 patches][chunk of 8 k][patch][16 B], the weights once at
// load time as [chunk][row][16 B] — so a K stage (4 chunks) of each operand is ONE contiguous block that lands in
// shared memory already in the canonical K-major no-swizzle UMMA layout (SBO = 128 B, LBO = rows*16 B).  (The first
// version described the row-major matrices to the TMA as (8 elements, rows, chunks) tensors: correct, but the copy
// engine then moved 16-byte pieces 4 KB apart.)  Three-stage mbarrier pipeline: 1 producer thread, 1 MMA-issuer
// thread, 8 epilogue warps.
constexpr int DN = 208;                       // dense1 outputs padded to a multiple of 16
constexpr int DM = 256;                       // patches per CTA
constexpr int DK_STAGE = 32;                  // K elements per stage (4 chunks of 16 B)
constexpr int D_STAGES = 3;
constexpr int D_A_BYTES = DM * DK_STAGE * 2;  // 16384
constexpr int D_W_BYTES = DN * DK_STAGE * 2;  // 13312
constexpr int D_STAGE = 2 * D_A_BYTES + 2 * D_W_BYTES;  // 59392
constexpr int D_SM_BAR = D_STAGES * D_STAGE;  // 178176
constexpr int D_SMEM = D_SM_BAR + 128;
constexpr int D_THREADS = 320;                // warps 0-3 and 6-9: epilogue of tile 0 / tile 1; warp 4: MMA issuer; warp 5: TMA producer

struct DenseArgs {
    const float *bd1, *d2, *bd2;  // (200), (200,20), (20)
    const __half *act3_hi, *act3_lo;   // tile-major A operands (conv3_tc_kernel)
    const __half *w_hi, *w_lo;         // dense1 weights, [256 chunks][208 rows][8] (prep_dense_weights_kernel)
    float *feat;
    int P, feat_stride, feat_col0;
    // frame mode: packed order is [F,3,K]; row p -> feat[(f*K+k)*60 + s*20]
    int frame_mode, K;
    long long *timeline;  // debug: [grid][8] clock64 stamps or null
};

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t sdst, const void *gsrc, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst),
                 "l"(gsrc), "r"(bytes), "r"(mbar)
                 : "memory");
}

__global__ void __launch_bounds__(D_THREADS, 1)
dense_tc_kernel(const DenseArgs a)
{
    extern __shared__ __align__(128) unsigned char dsm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(dsm + D_SM_BAR), *empty = full + D_STAGES, *accum = full + 2 * D_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dsm + D_SM_BAR + 96);
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int row0 = blockIdx.x * DM;
    auto stamp = [&](int slot) { if (a.timeline) a.timeline[(size_t)blockIdx.x * 8 + slot] = clock64(); };
    if (tid == 0) stamp(0);
    if (warp == 4) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int s = 0; s < D_STAGES; ++s) {
            umma::mbar_init(&full[s], 1);
            umma::mbar_init(&empty[s], 1);
        }
        umma::mbar_init(accum, 1);
        umma::fence_mbar_init();
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t sbase = umma::smem_u32(dsm);
    constexpr int NKT = 2048 / DK_STAGE;

    if (warp == 5) {
        // ===== TMA producer =====
        if (umma::elect_one()) {
            for (int kt = 0; kt < NKT; ++kt) {
                const int s = kt % D_STAGES;
                if (kt >= D_STAGES) umma::mbar_wait(&empty[s], (uint32_t)(((kt / D_STAGES) - 1) & 1));
                const uint32_t st = sbase + s * D_STAGE, mb = umma::smem_u32(&full[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(D_STAGE) : "memory");
                // operands are stored tile-major ([tile][chunk][row][16 B]): the stage's 4 chunks of each operand are one
                // contiguous block that lands in shared memory already in the canonical K-major layout
                const size_t a_off = ((size_t)blockIdx.x * 256 + (size_t)kt * (DK_STAGE / 8)) * DM * 8;   // halves
                const size_t w_off = (size_t)kt * (DK_STAGE / 8) * DN * 8;
                bulk_g2s(st, a.act3_hi + a_off, D_A_BYTES, mb);
                bulk_g2s(st + D_A_BYTES, a.act3_lo + a_off, D_A_BYTES, mb);
                bulk_g2s(st + 2 * D_A_BYTES, a.w_hi + w_off, D_W_BYTES, mb);
                bulk_g2s(st + 2 * D_A_BYTES + D_W_BYTES, a.w_lo + w_off, D_W_BYTES, mb);
            }
        }
        __syncwarp();
    } else if (warp == 4) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma::idesc_f16_f32(128, DN);
        for (int kt = 0; kt < NKT; ++kt) {
            const int s = kt % D_STAGES;
            umma::mbar_wait(&full[s], (uint32_t)((kt / D_STAGES) & 1));
            umma::fence_after_thread_sync();
            if (umma::elect_one()) {
                const uint32_t st = sbase + s * D_STAGE;
#pragma unroll
                for (int j = 0; j < DK_STAGE / 16; ++j) {  // chunk stride = rows * 16 B
                    const uint64_t wh = umma::smem_desc(st + 2 * D_A_BYTES + j * 2 * (DN * 16), DN * 16, 128);
                    const uint64_t wl = umma::smem_desc(st + 2 * D_A_BYTES + D_W_BYTES + j * 2 * (DN * 16), DN * 16, 128);
                    const uint32_t acc = (kt | j) ? 1u : 0u;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {          // the two 128-row tiles of the CTA
                        const uint64_t ah = umma::smem_desc(st + j * 2 * (DM * 16) + t * 2048, DM * 16, 128);
                        const uint64_t al = umma::smem_desc(st + D_A_BYTES + j * 2 * (DM * 16) + t * 2048, DM * 16, 128);
                        umma::mma_f16(tbase + t * 256, ah, wh, idesc, acc);
                        umma::mma_f16(tbase + t * 256, al, wh, idesc, 1u);
                        umma::mma_f16(tbase + t * 256, ah, wl, idesc, 1u);
                    }
                }
                umma::commit(&empty[s]);
                if (kt == NKT - 1) umma::commit(accum);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: h = tanh(D + b) in registers -> dense2 (200 -> 20) + tanh; one warp group per 128-row tile =====
        const int et = warp >= 6 ? 1 : 0, q = warp & 3;           // tile, TMEM lane quarter (a warp may touch lanes 32*(warp%4)..)
        const int row = q * 32 + (tid & 31), eid = et * 128 + row;
        float *W2s = reinterpret_cast<float *>(dsm);              // [200][20] + bd2[20] (+pad) + bd1[208]; aliases the operand stages
        if (eid == 0) stamp(2);
        umma::mbar_wait(accum, 0);                               // all MMAs done: the stages are free
        umma::fence_after_thread_sync();
        if (eid == 0) stamp(3);
        for (int i = eid; i < 200 * 20; i += 256) W2s[i] = a.d2[i];
        if (eid < 20) W2s[4000 + eid] = a.bd2[eid];
        if (eid < 208) W2s[4032 + eid] = eid < 200 ? a.bd1[eid] : 0.0f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (eid == 0) stamp(4);
        float2 acc2[10];                                         // dense2 accumulators, two outputs per packed FFMA2
#pragma unroll
        for (int j = 0; j < 10; ++j) acc2[j] = make_float2(W2s[4000 + 2 * j], W2s[4001 + 2 * j]);
        const uint32_t trow = tbase + ((uint32_t)(q * 32) << 16) + et * 256;
#pragma unroll 1
        for (int c0 = 0; c0 < DN; c0 += 16) {
            uint32_t v0[16];
            umma::tmem_ld_x16(trow + c0, v0);
            umma::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (c0 + j >= 200) break;                        // columns 200..207 are padding
                const float h = fast_tanh(__uint_as_float(v0[j]) + W2s[4032 + c0 + j]);
                const float2 hh = make_float2(h, h);
                const float4 *w = reinterpret_cast<const float4 *>(W2s + (c0 + j) * 20);
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const float4 wv = w[u];
                    acc2[2 * u] = __ffma2_rn(hh, make_float2(wv.x, wv.y), acc2[2 * u]);
                    acc2[2 * u + 1] = __ffma2_rn(hh, make_float2(wv.z, wv.w), acc2[2 * u + 1]);
                }
            }
        }
        float acc[20];
#pragma unroll
        for (int j = 0; j < 10; ++j) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
        const int p = row0 + et * 128 + row;
        if (p < a.P) {
            float *o;
            if (a.frame_mode) {
                int k = p % a.K, sc = (p / a.K) % 3, f = p / (3 * a.K);
                o = a.feat + ((size_t)f * a.K + k) * 60 + sc * 20;
            } else {
                o = a.feat + (size_t)p * a.feat_stride + a.feat_col0;
            }
#pragma unroll
            for (int j = 0; j < 20; ++j) o[j] = tanhf(acc[j]);
        }
        if (eid == 0) stamp(5);
        umma::fence_before_thread_sync();
    }
    __syncthreads();
    umma::fence_after_thread_sync();
    if (warp == 4) umma::tmem_dealloc(tbase, 512);
}

// conv12 tables, computed once per weight set.
//   T1 [9 (dx,dy)][8 dz-patterns][8 ch]: partial conv1 sums  sum_{dz in pattern} k1[(dx,dy,dz)][ch]
//   BG [27 classes][16 ch]: conv2 + max-pool + tanh where the whole neighbourhood is background.  If every
//   conv1 cell a pooled conv2 output can see is empty, conv1+pool+tanh is tanh(b1) inside the volume and 0 in
//   the halo, so the pooled output only depends on whether its coordinate touches the low edge (class 0),
//   the high edge (2) or neither (1) on each axis.
__global__ void prep_conv12_tables_kernel(const float *k1, const float *b1, const float *k2, const float *b2, float *tables)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < 576) {
        const int j = e >> 6, pat = (e >> 3) & 7, c = e & 7;
        float s = 0.f;
        for (int dz = 0; dz < 3; ++dz)
            if ((pat >> dz) & 1) s += k1[(j * 3 + dz) * 8 + c];
        tables[e] = s;
    } else if (e < 576 + 27 * 16) {
        const int cls = (e - 576) >> 4, co = (e - 576) & 15;
        const int pc[3] = {cls / 9, (cls / 3) % 3, cls % 3};
        float bgv[8];
        for (int ci = 0; ci < 8; ++ci) bgv[ci] = tanhf(b1[ci]);
        float best = -3.0e38f;
        for (int sub = 0; sub < 8; ++sub) {
            // edge class of the sub-position on each axis: 0 = first cell (tap 0 reads the halo), 2 = last cell
            int ec[3];
            for (int ax = 0; ax < 3; ++ax) {
                const int sb = (sub >> (2 - ax)) & 1;
                ec[ax] = (pc[ax] == 0 && sb == 0) ? 0 : ((pc[ax] == 2 && sb == 1) ? 2 : 1);
            }
            float acc = 0.f;
            for (int t = 0; t < 27; ++t) {
                const int d[3] = {t / 9, (t / 3) % 3, t % 3};
                bool in = true;
                for (int ax = 0; ax < 3; ++ax) in = in && !(ec[ax] == 0 && d[ax] == 0) && !(ec[ax] == 2 && d[ax] == 2);
                if (!in) continue;
                for (int ci = 0; ci < 8; ++ci) acc = fmaf(bgv[ci], k2[(t * 8 + ci) * 16 + co], acc);
            }
            best = fmaxf(best, acc);
        }
        tables[e] = fast_tanh(best + b2[co]);
    }
}

// dense1 weights (2048,200) f32 -> split fp16 in the operand layout [chunk of 8 k (256)][row n (208)][8] (rows 200..207 zero)
__global__ void prep_dense_weights_kernel(const float *d1, __half *w_hi, __half *w_lo)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < DN * 2048; i += gridDim.x * blockDim.x) {
        int n = i / 2048, k = i % 2048;
        __half h = __float2half_rn(0.f), l = h;
        if (n < 200) umma::split_f16(d1[(size_t)k * 200 + n], h, l);
        const size_t o = ((size_t)(k >> 3) * DN + n) * 8 + (k & 7);
        w_hi[o] = h;
        w_lo[o] = l;
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
    const size_t Ppad = ((size_t)P + DM - 1) / DM * DM;
    int rc = caelo_reserve(ctx, ctx->enc_ws, Ppad * 2048 * 2 * 2 + (size_t)P * 1024 * 4);
    if (rc) return rc;
    __half *act3_hi = reinterpret_cast<__half *>(ctx->enc_ws.ptr);
    __half *act3_lo = act3_hi + Ppad * 2048;
    float *act2 = reinterpret_cast<float *>(act3_lo + Ppad * 2048);
    Conv12Args c;
    c.packed = packed; c.k1 = ctx->enc.k1; c.b1 = ctx->enc.b1; c.k2 = ctx->enc.k2; c.b2 = ctx->enc.b2;
    c.act2 = act2; c.P = P; c.K3 = frame_mode ? 3 * K : 0; c.timeline = ctx->dbg_timeline;
    c.tables = ctx->enc_c12_tables;
    {
        const char *e = getenv("CAELO_CONV12_SKIP_BG");   // debug switch for A/B timing; default on
        c.skip_bg = !(e && e[0] == '0');
        e = getenv("CAELO_CONV12_DBG");
        c.dbg = e ? atoi(e) : 0;
    }
    {
        const char *e = getenv("CAELO_CONV12_PAIR");      // switch for A/B timing: "0" = the one-patch-per-MMA kernel (M = 64)
        const bool one_patch = e && e[0] == '0';
        ProfScope ps_(ctx, one_patch ? "conv12_tc_kernel" : "conv12_pair_kernel", st);
        if (one_patch) {
            int grid = 2 * ctx->num_sms < P ? 2 * ctx->num_sms : P;
            conv12_tc_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(c);
        } else {
            int grid = ctx->num_sms < (P + 1) / 2 ? ctx->num_sms : (P + 1) / 2;   // persistent: one CTA per SM walks the patch PAIRS
            conv12_pair_kernel<<<grid, P2_THREADS, P2_SMEM, st>>>(c);
        }
    }
    CAELO_LAUNCH_CHECK(ctx);
    Conv3Args c3;
    c3.act2 = act2; c3.k3 = ctx->enc.k3; c3.b3 = ctx->enc.b3; c3.act3_hi = act3_hi; c3.act3_lo = act3_lo; c3.P = P;
    { const char *e = getenv("CAELO_CONV3_DBG"); c3.dbg = e ? atoi(e) : 0; }
    {
        const char *e8 = getenv("CAELO_CONV3_OCT");      // switch for A/B timing: "0" = the one-patch-per-MMA kernel (M = 64)
        const bool oct = !(e8 && e8[0] == '0');
        ProfScope ps_(ctx, oct ? "conv3_oct_kernel" : "conv3_tc_kernel", st);
        if (oct) {
            int grid3 = ctx->num_sms;                    // persistent: one CTA per SM walks the groups of eight patches
            if (grid3 > (P + 7) / 8) grid3 = (P + 7) / 8;
            conv3_oct_kernel<<<grid3, C8_THREADS, C8_SMEM, st>>>(c3);
        } else {
            // one CTA per SM with two operand buffers (two single-buffer CTAs per SM measured slower: 0.95 vs 0.89 ms)
            int grid3 = ctx->num_sms;
            if (grid3 > P) grid3 = P;
            conv3_tc_kernel<2><<<grid3, C3_THREADS, c3_smem(2), st>>>(c3);
        }
    }
    CAELO_LAUNCH_CHECK(ctx);
    DenseArgs d;
    d.bd1 = ctx->enc.bd1; d.d2 = ctx->enc.d2; d.bd2 = ctx->enc.bd2;
    d.act3_hi = act3_hi; d.act3_lo = act3_lo; d.w_hi = ctx->enc_w1t_hi; d.w_lo = ctx->enc_w1t_lo;
    d.feat = feat; d.P = P; d.feat_stride = feat_stride; d.feat_col0 = feat_col0;
    d.frame_mode = frame_mode; d.K = K; d.timeline = ctx->dbg_timeline ? ctx->dbg_timeline + (size_t)2 * ctx->num_sms * 64 * 16 : nullptr;
    { ProfScope ps_(ctx, "dense_tc_kernel", st); dense_tc_kernel<<<(unsigned)(Ppad / DM), D_THREADS, D_SMEM, st>>>(d); }
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}

}  // namespace

int caelo_encoder_init(caelo_ctx *ctx)
{
    CAELO_CUDA(ctx, cudaFuncSetAttribute(conv12_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(conv12_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(conv3_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, c3_smem(2)));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(conv3_oct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C8_SMEM));
    CAELO_CUDA(ctx, cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM));
    return CAELO_OK;
}

// called from caelo_set_encoder_weights: derived operand forms of the weights
int caelo_encoder_prepare(caelo_ctx *ctx)
{
    if (!ctx->enc_w1t_hi) {
        CAELO_CUDA(ctx, cudaMalloc(&ctx->enc_w1t_hi, (size_t)DN * 2048 * 2 * 2));
        ctx->enc_w1t_lo = ctx->enc_w1t_hi + (size_t)DN * 2048;
    }
    prep_dense_weights_kernel<<<256, 256>>>(ctx->enc.d1, ctx->enc_w1t_hi, ctx->enc_w1t_lo);
    CAELO_LAUNCH_CHECK(ctx);
    if (!ctx->enc_c12_tables) CAELO_CUDA(ctx, cudaMalloc(&ctx->enc_c12_tables, (576 + 27 * 16) * 4));
    prep_conv12_tables_kernel<<<4, 256>>>(ctx->enc.k1, ctx->enc.b1, ctx->enc.k2, ctx->enc.b2, ctx->enc_c12_tables);
    CAELO_LAUNCH_CHECK(ctx);
    CAELO_CUDA(ctx, cudaDeviceSynchronize());
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
    if (status) CAELO_CUDA(ctx, caelo_fill_async(status, 0, 4, st));
    long long nwords = (long long)P * 128;
    long long blocks = (nwords * 32 + 255) / 256;
    if (blocks > (long long)ctx->num_sms * 32) blocks = (long long)ctx->num_sms * 32;
    { ProfScope ps_(ctx, "pack_kernel", st); pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(patches, packed, nwords, status); }
    CAELO_LAUNCH_CHECK(ctx);
    return run_encoder(ctx, packed, P, feat, 20, 0, 0, 1, st);
}
