// patches.cu — a6: GetPatchesList (reference Voxel.py:177-216) for sm_100a.
//
// The reference fits an sklearn kNN index on each of the three occupied-voxel lists, asks for
// the 496 nearest voxels of every key voxel, keeps those inside the [-8,8)^3 cube and scatters
// 1.0 into a 16^3 patch using NEGATIVE indices (offset o lands at index o mod 16).  Here:
//   1. brick_insert_kernel: every occupied voxel sets one bit of a 4x4x4 "brick" (64-bit mask)
//      kept in an open-addressing hash table keyed by the brick coordinate (per frame, scale);
//      and the thread that claims a brick slot also sets the brick's bit in a second table keyed by
//      the 4x4x4-brick "super brick" (16^3 voxels);
//   2. gather_kernel: one warp per (frame, scale, keypoint) looks up the <=27 super bricks that meet
//      the ball d^2 <= 192 around the key voxel, lists their non-empty bricks (most of the <=8^3
//      bricks around a key voxel are empty: one probe per super brick instead of one per brick),
//      probes those, counts the ball and sets the cube bits of a 512-byte bit-packed patch in
//      shared memory; only if the ball holds more than 496 voxels (the kNN cut can bite) it
//      re-runs with the exact rank rule (d^2, then x,y,z).
// Integer-exact; HBM traffic = voxel lists in + 512 B per patch out.
#include "common.cuh"
#include "voxel_math.cuh"

namespace {

constexpr unsigned long long EMPTY = ~0ull;
constexpr int kWarps = 8;
constexpr int NNB = 496;  // n_neighbors, Voxel.py:182

// Open-addressing table of 16-byte slots {key, ~mask}: key and occupancy word share one 32-byte sector (a probe is
// one memory transaction), and the whole table is cleared by ONE memset to 0xFF (empty key = all ones, the mask is
// stored inverted: a set voxel/brick bit is a CLEARED bit).
struct Table {
    unsigned long long *slots;   // [2 * capacity]: slots[2i] = key, slots[2i+1] = ~mask
    unsigned cap_mask;           // capacity - 1 (power of two)
};
__device__ __forceinline__ unsigned long long *key_of(const Table &t, unsigned slot) { return t.slots + 2 * (size_t)slot; }
__device__ __forceinline__ unsigned long long *inv_mask_of(const Table &t, unsigned slot) { return t.slots + 2 * (size_t)slot + 1; }

__device__ __forceinline__ unsigned hash64(unsigned long long k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned)k;
}

__device__ __forceinline__ unsigned long long brick_key(int bx, int by, int bz)
{
    return (unsigned long long)(bx + 8192) | ((unsigned long long)(by + 8192) << 14) |
           ((unsigned long long)(bz + 8192) << 28);
}

// the brick (bx,by,bz) exists: set its bit in the super-brick table (called once per brick, by the claimer of its slot)
__device__ __forceinline__ void super_set(const Table &t, int bx, int by, int bz)
{
    const unsigned long long key = brick_key(bx >> 2, by >> 2, bz >> 2);
    const unsigned long long bit = 1ull << (((bx & 3) * 4 + (by & 3)) * 4 + (bz & 3));
    unsigned slot = hash64(key) & t.cap_mask;
    while (true) {
        unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(key_of(t, slot));
        if (k == EMPTY) {
            k = atomicCAS(key_of(t, slot), EMPTY, key);
            if (k == EMPTY) k = key;
        }
        if (k == key) {
            atomicAnd(inv_mask_of(t, slot), ~bit);
            return;
        }
        slot = (slot + 1) & t.cap_mask;
    }
}

struct BuildArgs {
    const int16_t *vox;          // all lists concatenated, rows of 3
    const long long *offsets;    // dev [F*3+1]
    const Table *tables;         // dev [2*F*3]: brick tables, then super-brick tables
    int nlists;
};

__global__ void __launch_bounds__(256) brick_insert_kernel(const BuildArgs a)
{
    const int list = blockIdx.y, lane = threadIdx.x & 31;
    const long long beg = a.offsets[list], end = a.offsets[list + 1];
    const Table t = a.tables[list];
    // Consecutive voxels of a list are neighbours in space (scan order), so several lanes of a warp usually hit the
    // same brick — and atomics on one address serialise in L2: lanes with equal keys merge their bits first and
    // only the group leader touches the table (warp-uniform trip count for the match / reduce).
    for (long long i0 = beg + (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < end;
         i0 += (long long)gridDim.x * blockDim.x) {
        const long long i = i0 + lane;
        const bool valid = i < end;
        int x = 0, y = 0, z = 0;
        if (valid) { x = a.vox[i * 3 + 0]; y = a.vox[i * 3 + 1]; z = a.vox[i * 3 + 2]; }
        const unsigned long long key = valid ? brick_key(x >> 2, y >> 2, z >> 2) : EMPTY;
        const unsigned long long bit = valid ? 1ull << (((x & 3) * 4 + (y & 3)) * 4 + (z & 3)) : 0ull;
        const unsigned grp = __match_any_sync(0xffffffffu, key);
        const unsigned lo = __reduce_or_sync(grp, (unsigned)bit), hi = __reduce_or_sync(grp, (unsigned)(bit >> 32));
        if (!valid || lane != __ffs(grp) - 1) continue;
        const unsigned long long bits = ((unsigned long long)hi << 32) | lo;
        unsigned slot = hash64(key) & t.cap_mask;
        while (true) {
            // most voxels fall into a brick that is already claimed: a plain load sees it, no CAS needed
            unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(key_of(t, slot));
            bool claimed = false;
            if (old == EMPTY) {
                old = atomicCAS(key_of(t, slot), EMPTY, key);
                claimed = old == EMPTY;
            }
            if (claimed || old == key) {
                atomicAnd(inv_mask_of(t, slot), ~bits);
                if (claimed) super_set(a.tables[a.nlists + list], x >> 2, y >> 2, z >> 2);
                break;
            }
            slot = (slot + 1) & t.cap_mask;
        }
    }
}

// Fused f2+a6 front end: bricks straight from the raw scans (no voxel lists).  The voxel SET is the one
// Voxelization (Voxel.py:100-173) produces — same arithmetic (voxel_math.cuh) — and the list ORDER never
// matters to GetPatchesList.  nvox[f*3+s] counts the distinct voxels (= len(AllVoxels_s)) for the
// sklearn "n_neighbors <= n_samples_fit" check.
struct ScanBuildArgs {
    const float *pts;            // rows of 4
    const long long *offsets;    // dev [F+1]
    const Table *tables;         // dev [2*F*3]: brick tables, then super-brick tables
    int nlists;                  // F*3
    int *nvox;                   // dev [F*3]
    int *status;                 // dev [F] or null
};

// Insert of one voxel, split in two so that a thread can have the first probes of all three scales in flight at
// once (the tables are far larger than L2: a probe is a DRAM round trip).  brick_probe issues ONE 16-byte load of
// the home slot {key, ~mask}; brick_set finishes the insert from it and returns true iff this call set a bit that
// was clear (a voxel seen for the first time).  Most points of a scan fall into a voxel that is already recorded:
// the loaded slot shows it and no atomic is issued.  (A stale load can only show LESS than the truth — bits are
// only ever set — and then the atomic's own return value decides.)
struct Probe {
    unsigned long long key, bit;
    unsigned slot;
    ulonglong2 kv;
};

__device__ __forceinline__ Probe brick_probe(const Table &t, int x, int y, int z)
{
    Probe p;
    p.key = brick_key(x >> 2, y >> 2, z >> 2);
    p.bit = 1ull << (((x & 3) * 4 + (y & 3)) * 4 + (z & 3));
    p.slot = hash64(p.key) & t.cap_mask;
    p.kv = __ldcg(reinterpret_cast<const ulonglong2 *>(key_of(t, p.slot)));
    return p;
}

__device__ __forceinline__ bool brick_set(const Table &t, const Table &ts, int x, int y, int z, const Probe &p)
{
    const unsigned long long key = p.key, bit = p.bit;
    unsigned slot = p.slot;
    ulonglong2 kv = p.kv;
    while (true) {
        unsigned long long k = kv.x;
        bool fresh = false;                      // this thread just claimed the slot: its mask is still empty
        if (k == EMPTY) {
            k = atomicCAS(key_of(t, slot), EMPTY, key);
            if (k == EMPTY) {
                k = key;
                fresh = true;
                super_set(ts, x >> 2, y >> 2, z >> 2);
            } else {
                kv.y = ~0ull;                    // someone else took it: nothing known about its mask
            }
        }
        if (k == key) {
            if (!fresh && !(kv.y & bit)) return false;   // already set
            return (atomicAnd(inv_mask_of(t, slot), ~bit) & bit) != 0ull;
        }
        slot = (slot + 1) & t.cap_mask;
        kv = __ldcg(reinterpret_cast<const ulonglong2 *>(key_of(t, slot)));
    }
}

__global__ void __launch_bounds__(256) scan_brick_insert_kernel(const ScanBuildArgs a)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const long long beg = a.offsets[f], n = a.offsets[f + 1] - beg;
    const float4 *p4 = reinterpret_cast<const float4 *>(a.pts) + beg;
    const Table t0 = a.tables[f * 3], t1 = a.tables[f * 3 + 1], t2 = a.tables[f * 3 + 2];
    const Table s0 = a.tables[a.nlists + f * 3], s1 = a.tables[a.nlists + f * 3 + 1], s2 = a.tables[a.nlists + f * 3 + 2];
    int bad = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // warp-uniform trip count: the ballots below need every lane
    for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
        const long long i = i0 + lane;
        bool n0 = false, n1 = false, n2 = false;
        if (i < n) {
            const float4 p = p4[i];
            VoxelOfPoint v;
            const int st = voxel_of_point(p.x, p.y, p.z, v);
            bad += st < 0;
            if (st > 0) {
                const Probe p0 = brick_probe(t0, v.g0[0], v.g0[1], v.g0[2]), p1 = brick_probe(t1, v.g1[0], v.g1[1], v.g1[2]),
                            p2 = brick_probe(t2, v.g2[0], v.g2[1], v.g2[2]);     // three DRAM round trips side by side
                n0 = brick_set(t0, s0, v.g0[0], v.g0[1], v.g0[2], p0);
                n1 = brick_set(t1, s1, v.g1[0], v.g1[1], v.g1[2], p1);
                n2 = brick_set(t2, s2, v.g2[0], v.g2[1], v.g2[2], p2);
            }
        }
        // warp-aggregated voxel counters (one atomic per warp and scale instead of one per new voxel)
        const unsigned m0 = __ballot_sync(0xffffffffu, n0), m1 = __ballot_sync(0xffffffffu, n1),
                       m2 = __ballot_sync(0xffffffffu, n2);
        if (lane == 0) {
            if (m0) atomicAdd(a.nvox + f * 3, __popc(m0));
            if (m1) atomicAdd(a.nvox + f * 3 + 1, __popc(m1));
            if (m2) atomicAdd(a.nvox + f * 3 + 2, __popc(m2));
        }
    }
    if (bad && a.status) atomicAdd(a.status + f, bad);
}

__device__ __forceinline__ unsigned long long brick_lookup(const Table &t, unsigned long long key)
{
    unsigned slot = hash64(key) & t.cap_mask;
    while (true) {
        const ulonglong2 kv = *reinterpret_cast<const ulonglong2 *>(key_of(t, slot));   // one 16-byte load: key, ~mask
        if (kv.x == key) return ~kv.y;
        if (kv.x == EMPTY) return 0ull;
        slot = (slot + 1) & t.cap_mask;
    }
}

struct GatherArgs {
    const void *kpts;            // [F,K,3] f32 or f64
    const int *n_kpts;           // [F] or null
    const Table *tables;         // [2*F*3]: brick tables, then super-brick tables
    unsigned *packed;            // [F,3,K,128]
    unsigned char *trunc;        // [F,3,K] or null
    const int *nvox;             // [F*3] distinct voxels per list, or null (checked on the host)
    int *few;                    // [F]: |= 2 where a list has < 496 voxels (sklearn raises), with nvox
    double vis[3];               // VisibleLength/Width/Height (Voxel.py:50-52)
    double vsize[3];             // VoxelSizes (Voxel.py:31)
    int kpts_f64, F, K;
};

// List the non-empty bricks that meet [-13,13]^3 around the key voxel: lanes 0..26 look up the <=3^3 super
// bricks, restrict their child masks to the brick range of the box and write the brick coordinates (relative to
// the box corner, 3 bits per axis) into `list` at warp-scanned offsets.  Returns the count (<= 512).
__device__ __forceinline__ int list_ball_bricks(const Table &ts, int kx, int ky, int kz, int lane, unsigned short *list,
                                                int &bx0, int &by0, int &bz0)
{
    bx0 = (kx - 13) >> 2; by0 = (ky - 13) >> 2; bz0 = (kz - 13) >> 2;
    const int bx1 = (kx + 13) >> 2, by1 = (ky + 13) >> 2, bz1 = (kz + 13) >> 2;
    const int sx0 = bx0 >> 2, sy0 = by0 >> 2, sz0 = bz0 >> 2;
    const int nsy = (by1 >> 2) - sy0 + 1, nsz = (bz1 >> 2) - sz0 + 1;
    const int ns = ((bx1 >> 2) - sx0 + 1) * nsy * nsz;
    unsigned long long m = 0ull;
    int sbx = 0, sby = 0, sbz = 0;
    if (lane < ns) {
        sbz = sz0 + lane % nsz; sby = sy0 + (lane / nsz) % nsy; sbx = sx0 + lane / (nsz * nsy);
        if (sbx >= -8192 && sby >= -8192 && sbz >= -8192) m = brick_lookup(ts, brick_key(sbx, sby, sbz));
        if (m) {
            // children inside the box: c in [max(b0 - 4 sb, 0), min(b1 - 4 sb, 3)] on each axis
            auto range = [](int lo, int hi) { lo = lo < 0 ? 0 : lo; hi = hi > 3 ? 3 : hi; return (0xFu >> (3 - hi)) & (0xFu << lo); };
            const unsigned zr = range(bz0 - 4 * sbz, bz1 - 4 * sbz), yr = range(by0 - 4 * sby, by1 - 4 * sby),
                           xr = range(bx0 - 4 * sbx, bx1 - 4 * sbx);
            unsigned yz = 0;  // 16 bits: (cy,cz)
#pragma unroll
            for (int c = 0; c < 4; ++c) yz |= ((yr >> c) & 1u) ? (zr << (4 * c)) : 0u;
            unsigned long long box = 0ull;
#pragma unroll
            for (int c = 0; c < 4; ++c) box |= ((xr >> c) & 1u) ? ((unsigned long long)yz << (16 * c)) : 0ull;
            m &= box;
        }
    }
    const int cnt = __popcll(m);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int pos = incl - cnt;
    while (m) {
        const int bit = __ffsll((long long)m) - 1;
        m &= m - 1;
        const int bx = 4 * sbx + (bit >> 4), by = 4 * sby + ((bit >> 2) & 3), bz = 4 * sbz + (bit & 3);
        list[pos++] = (unsigned short)(((bx - bx0) << 6) | ((by - by0) << 3) | (bz - bz0));
    }
    __syncwarp();
    return total;
}

// visit every occupied voxel of the listed bricks; f(dx,dy,dz) relative to the key voxel
template <class Fn>
__device__ __forceinline__ void for_ball_voxels(const Table &t, const unsigned short *list, int nlist, int bx0, int by0,
                                                int bz0, int kx, int ky, int kz, int lane, Fn f)
{
    for (int i = lane; i < nlist; i += 32) {
        const unsigned e = list[i];
        const int bx = bx0 + (int)(e >> 6), by = by0 + (int)((e >> 3) & 7), bz = bz0 + (int)(e & 7);
        unsigned long long m = brick_lookup(t, brick_key(bx, by, bz));
        while (m) {
            int bit = __ffsll((long long)m) - 1;
            m &= m - 1;
            int dx = bx * 4 + (bit >> 4) - kx, dy = by * 4 + ((bit >> 2) & 3) - ky,
                dz = bz * 4 + (bit & 3) - kz;
            f(dx, dy, dz);
        }
    }
}

__device__ __forceinline__ bool in_cube(int dx, int dy, int dz)
{
    return dx >= -8 && dx < 8 && dy >= -8 && dy < 8 && dz >= -8 && dz < 8;
}

__device__ __forceinline__ void set_patch_bit(unsigned *patch, int dx, int dy, int dz)
{
    int idx = (((dx & 15) * 16) + (dy & 15)) * 16 + (dz & 15);  // o mod 16: the negative-index roll
    atomicOr(patch + (idx >> 5), 1u << (idx & 31));
}

__global__ void __launch_bounds__(kWarps * 32) gather_kernel(const GatherArgs a)
{
    __shared__ unsigned s_patch[kWarps][128];
    __shared__ int s_hist[kWarps][196];
    __shared__ unsigned s_list[kWarps][256];
    __shared__ unsigned short s_bricks[kWarps][512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total = (long long)a.F * 3 * a.K;
    unsigned *patch = s_patch[warp];
    for (long long w = (long long)blockIdx.x * kWarps + warp; w < total;
         w += (long long)gridDim.x * kWarps) {
        const int k = (int)(w % a.K);
        const int s = (int)((w / a.K) % 3);
        const int f = (int)(w / (3LL * a.K));
        for (int i = lane; i < 128; i += 32) patch[i] = 0u;
        __syncwarp();
        bool live = a.n_kpts == nullptr || k < a.n_kpts[f];
        if (a.nvox && a.nvox[f * 3 + s] < NNB) {   // n_neighbors <= n_samples_fit (Voxel.py:195)
            if (lane == 0 && k == 0) atomicOr(a.few + f, 0x40000000);
            live = false;
        }
        bool flagged = false;
        if (live) {
            double p[3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                p[c] = a.kpts_f64 ? reinterpret_cast<const double *>(a.kpts)[((size_t)f * a.K + k) * 3 + c]
                                  : (double)reinterpret_cast<const float *>(a.kpts)[((size_t)f * a.K + k) * 3 + c];
            // KeyVoxels = int32((Pts + Visible) / VoxelSizes[s])  — float64, truncation (Voxel.py:185,193)
            const double vsz = s == 0 ? a.vsize[0] : (s == 1 ? a.vsize[1] : a.vsize[2]);  // no dynamic param indexing
            const int kx = (int)__ddiv_rn(__dadd_rn(p[0], a.vis[0]), vsz);
            const int ky = (int)__ddiv_rn(__dadd_rn(p[1], a.vis[1]), vsz);
            const int kz = (int)__ddiv_rn(__dadd_rn(p[2], a.vis[2]), vsz);
            const Table t = a.tables[f * 3 + s];
            int bx0, by0, bz0;
            const unsigned short *bricks = s_bricks[warp];
            const int nb = list_ball_bricks(a.tables[a.F * 3 + f * 3 + s], kx, ky, kz, lane, s_bricks[warp], bx0, by0, bz0);
            int ball = 0;
            for_ball_voxels(t, bricks, nb, bx0, by0, bz0, kx, ky, kz, lane, [&](int dx, int dy, int dz) {
                int d2 = dx * dx + dy * dy + dz * dz;
                if (d2 <= 192) {
                    ++ball;
                    if (in_cube(dx, dy, dz)) set_patch_bit(patch, dx, dy, dz);
                }
            });
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ball += __shfl_xor_sync(0xffffffffu, ball, o);
            __syncwarp();
            if (ball > NNB) {
                // the 496-NN cut can bite: keep cube voxels whose rank by (d2, x, y, z) is < 496
                flagged = true;
                int *hist = s_hist[warp];
                unsigned *list = s_list[warp];
                for (int i = lane; i < 196; i += 32) hist[i] = 0;
                for (int i = lane; i < 128; i += 32) patch[i] = 0u;
                __syncwarp();
                for_ball_voxels(t, bricks, nb, bx0, by0, bz0, kx, ky, kz, lane, [&](int dx, int dy, int dz) {
                    int d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 <= 192) atomicAdd(hist + d2, 1);
                });
                __syncwarp();
                // cut value q: smallest d2 with cumulative count >= 496 (every lane computes it)
                int q = 0, before = 0;
                for (; q <= 192; ++q) {
                    if (before + hist[q] >= NNB) break;
                    before += hist[q];
                }
                const int room = NNB - before;  // how many voxels at d2 == q survive
                if (lane == 0) hist[193] = 0;
                __syncwarp();
                for_ball_voxels(t, bricks, nb, bx0, by0, bz0, kx, ky, kz, lane, [&](int dx, int dy, int dz) {
                    int d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 < q) {
                        if (in_cube(dx, dy, dz)) set_patch_bit(patch, dx, dy, dz);
                    } else if (d2 == q) {
                        int pos = atomicAdd(hist + 193, 1);
                        if (pos < 256) list[pos] = ((unsigned)(dx + 16) << 12) | ((unsigned)(dy + 16) << 6) | (unsigned)(dz + 16);
                    }
                });
                __syncwarp();
                int m = hist[193];
                if (m > 256) m = 256;  // r3(n) <= 168 for n <= 192: cannot happen
                for (int i = lane; i < m; i += 32) {
                    unsigned me = list[i];
                    int rank = 0;
                    for (int j = 0; j < m; ++j) rank += list[j] < me;
                    if (rank < room) {
                        int dx = (int)(me >> 12) - 16, dy = (int)((me >> 6) & 63) - 16, dz = (int)(me & 63) - 16;
                        if (in_cube(dx, dy, dz)) set_patch_bit(patch, dx, dy, dz);
                    }
                }
                __syncwarp();
            }
        }
        unsigned *out = a.packed + (((size_t)f * 3 + s) * a.K + k) * 128;
        for (int i = lane; i < 128; i += 32) out[i] = patch[i];
        if (a.trunc && lane == 0) a.trunc[((size_t)f * 3 + s) * a.K + k] = flagged ? 1 : 0;
        __syncwarp();
    }
}

__global__ void unpack_kernel(const unsigned *__restrict__ packed, float *__restrict__ out, long long nwords)
{
    // one thread per packed word -> 32 floats (coalesced 128 B per thread pair of float4 x8)
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
         w += (long long)gridDim.x * blockDim.x) {
        unsigned v = packed[w];
        float4 *o = reinterpret_cast<float4 *>(out + w * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o[j] = make_float4((v >> (4 * j)) & 1u ? 1.f : 0.f, (v >> (4 * j + 1)) & 1u ? 1.f : 0.f,
                               (v >> (4 * j + 2)) & 1u ? 1.f : 0.f, (v >> (4 * j + 3)) & 1u ? 1.f : 0.f);
    }
}

}  // namespace

// Hash-table geometry + upload: caps[l] slots (power of two) for list l; `offsets` (n_off int64) is staged next
// to the Table array.  All tables are cleared on the stream.
static int setup_tables(caelo_ctx *ctx, int nl, const size_t *caps, const int64_t *offsets, int n_off, Table **d_tables,
                        long long **d_off, cudaStream_t st)
{
    size_t total_slots = 0;
    for (int l = 0; l < nl; ++l) total_slots += caps[l];
    const size_t head = ((size_t)nl * sizeof(Table) + (size_t)n_off * 8 + 255) / 256 * 256;
    int rc = caelo_reserve(ctx, ctx->bricks, head + total_slots * 16);
    if (rc) return rc;
    char *base = reinterpret_cast<char *>(ctx->bricks.ptr);
    *d_tables = reinterpret_cast<Table *>(base);
    *d_off = reinterpret_cast<long long *>(base + (size_t)nl * sizeof(Table));
    unsigned long long *d_slots = reinterpret_cast<unsigned long long *>(base + head);
    // tables + offsets go through a ring of pinned staging slots: no stream synchronisation
    const size_t stage_bytes = (size_t)nl * sizeof(Table) + (size_t)n_off * 8;
    void *h_stage = nullptr;
    cudaEvent_t ev;
    rc = caelo_stage_acquire(ctx, stage_bytes, &h_stage, &ev);
    if (rc) return rc;
    Table *h_tables = reinterpret_cast<Table *>(h_stage);
    size_t cur = 0;
    for (int l = 0; l < nl; ++l) {
        h_tables[l].slots = d_slots + 2 * cur;
        h_tables[l].cap_mask = (unsigned)(caps[l] - 1);
        cur += caps[l];
    }
    memcpy(reinterpret_cast<char *>(h_stage) + (size_t)nl * sizeof(Table), offsets, (size_t)n_off * 8);
    CAELO_CUDA(ctx, caelo_stage_copy_async(*d_tables, h_stage, stage_bytes, st));
    CAELO_CUDA(ctx, cudaEventRecord(ev, st));
    CAELO_CUDA(ctx, caelo_fill_async(d_slots, 0xFF, total_slots * 16, st));
    return CAELO_OK;
}

#ifndef CAELO_BRICK_SLOTS_NUM
#define CAELO_BRICK_SLOTS_NUM 4      // brick-table slots per voxel, in halves (2 = one slot per voxel, 4 = two)
#endif
static size_t pow2_above(size_t n)
{
    size_t cap = 1024;
    while (cap < n) cap <<= 1;
    return cap;
}

static int launch_gather(caelo_ctx *ctx, const void *kpts, int kpts_f64, const int32_t *n_kpts, int F, int K,
                         const Table *d_tables, const int *nvox, int *few, uint32_t *packed, float *patches_f32,
                         uint8_t *trunc, cudaStream_t st)
{
    GatherArgs g;
    g.kpts = kpts; g.n_kpts = n_kpts; g.tables = d_tables; g.packed = packed; g.trunc = trunc;
    g.nvox = nvox; g.few = few;
    // Voxel.py:40-52: nBlocksL = int(200/1.28) = 156, nBlocksH = int(30/1.28) = 23; Visible* = n/2*1.28
    const double brs = 1.28;
    g.vis[0] = 156 / 2.0 * brs; g.vis[1] = 156 / 2.0 * brs; g.vis[2] = 23 / 2.0 * brs;
    const double vs = 0.02;
    g.vsize[0] = vs; g.vsize[1] = vs * 8; g.vsize[2] = vs * 32;
    g.kpts_f64 = kpts_f64; g.F = F; g.K = K;
    long long warps = (long long)F * 3 * K;
    long long blocks = (warps + kWarps - 1) / kWarps;
    long long maxb = (long long)ctx->num_sms * 16;
    if (blocks > maxb) blocks = maxb;
    { ProfScope ps_(ctx, "gather_kernel", st); gather_kernel<<<(unsigned)blocks, kWarps * 32, 0, st>>>(g); }
    CAELO_LAUNCH_CHECK(ctx);

    if (patches_f32) {
        long long nwords = (long long)F * 3 * K * 128;
        long long ub = (nwords + 255) / 256;
        if (ub > (long long)ctx->num_sms * 32) ub = (long long)ctx->num_sms * 32;
        { ProfScope ps_(ctx, "unpack_kernel", st); unpack_kernel<<<(unsigned)ub, 256, 0, st>>>(packed, patches_f32, nwords); }
        CAELO_LAUNCH_CHECK(ctx);
    }
    return CAELO_OK;
}

// ---- a6 in two steps: the occupancy index of a batch does not depend on its key points, so a caller can build it
//      on a second stream while the key points are still being selected -----------------------------------------
extern "C" int caelo_bricks_build(caelo_ctx *ctx, const int16_t *vox, const int64_t *vox_offsets, int F, void *stream)
{
    if (!ctx || !vox || !vox_offsets || F <= 0) return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int nl = F * 3;
    std::vector<size_t> caps(2 * nl);
    long long maxlen = 0;
    for (int l = 0; l < nl; ++l) {
        long long len = vox_offsets[l + 1] - vox_offsets[l];
        if (len < NNB) return CAELO_ERR_TOO_FEW_VOXELS;  // sklearn: n_neighbors <= n_samples_fit
        if (len > maxlen) maxlen = len;
        // bricks <= voxels, so len + 1 slots could never fill up either — measured (round 2): 0.263 vs 0.229 ms for the insert,
        // 0.96 vs 0.86 ms for the from-scans build + gather: the longer probe chains cost more than the better L2 hit rate
        // of the smaller table gives.  The build is bound by the RATE of random DRAM accesses (~12 M random 32-byte sectors
        // per batch in 0.6 ms = 20 G/s at 0.6 TB/s), not by latency: inserting a point's three voxels side by side and
        // merging the lanes of a warp that hit one brick changed nothing.
        caps[l] = pow2_above((size_t)len * CAELO_BRICK_SLOTS_NUM / 2 + 1);
        caps[nl + l] = pow2_above((size_t)len + 2);      // super bricks <= bricks <= voxels
    }
    Table *d_tables;
    long long *d_off;
    int rc = setup_tables(ctx, 2 * nl, caps.data(), vox_offsets, nl + 1, &d_tables, &d_off, st);
    if (rc) return rc;
    BuildArgs b;
    b.vox = vox; b.offsets = d_off; b.tables = d_tables; b.nlists = nl;
    int bx = (int)((maxlen + 255) / 256);   // one voxel per thread: the kernel is latency-bound, parallelism is what it needs
    if (bx > 1024) bx = 1024;
    { ProfScope ps_(ctx, "brick_insert_kernel", st); brick_insert_kernel<<<dim3(bx, nl), 256, 0, st>>>(b); }
    CAELO_LAUNCH_CHECK(ctx);
    ctx->bricks_tables = d_tables;
    ctx->bricks_frames = F;
    return CAELO_OK;
}

extern "C" int caelo_bricks_build_scans(caelo_ctx *ctx, const float *pts, const int64_t *pts_offsets, int F, int32_t *nvox,
                                        int32_t *status, void *stream)
{
    if (!ctx || !pts || !pts_offsets || !nvox || !status || F <= 0) return CAELO_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int nl = F * 3;
    std::vector<size_t> caps(2 * nl);
    long long maxn = 0;
    for (int f = 0; f < F; ++f) {
        long long n = pts_offsets[f + 1] - pts_offsets[f];
        if (n < 0) return CAELO_ERR_ARG;
        if (n > maxn) maxn = n;
        // a brick holds >= 1 voxel and a voxel >= 1 point: n bounds every table; the 64 cm grid has
        // 78*78*12 = 73,008 bricks at most
        caps[f * 3] = pow2_above((size_t)n * CAELO_BRICK_SLOTS_NUM / 2 + 2);
        caps[f * 3 + 1] = pow2_above((size_t)n + 2);
        caps[f * 3 + 2] = pow2_above(((size_t)n < 73008 ? (size_t)n : 73008) * 2 + 2);
        // super bricks (4^3 bricks): never more than bricks; the 64 cm grid has 20*20*3 = 1,200 of them
        caps[nl + f * 3] = pow2_above((size_t)n + 2);
        caps[nl + f * 3 + 1] = pow2_above((size_t)n + 2);
        caps[nl + f * 3 + 2] = pow2_above(((size_t)n < 1200 ? (size_t)n : 1200) * 2 + 2);
    }
    Table *d_tables;
    long long *d_off;
    int rc = setup_tables(ctx, 2 * nl, caps.data(), pts_offsets, F + 1, &d_tables, &d_off, st);
    if (rc) return rc;
    CAELO_CUDA(ctx, caelo_fill_async(nvox, 0, (size_t)nl * 4, st));
    CAELO_CUDA(ctx, caelo_fill_async(status, 0, (size_t)F * 4, st));
    ScanBuildArgs b;
    b.pts = pts; b.offsets = d_off; b.tables = d_tables; b.nlists = nl; b.nvox = nvox; b.status = status;
    int bx = (int)((maxn + 255) / 256);     // one point per thread
    if (bx > 1024) bx = 1024;
    if (bx < 1) bx = 1;
    { ProfScope ps_(ctx, "scan_brick_insert_kernel", st); scan_brick_insert_kernel<<<dim3(bx, F), 256, 0, st>>>(b); }
    CAELO_LAUNCH_CHECK(ctx);
    ctx->bricks_tables = d_tables;
    ctx->bricks_frames = F;
    return CAELO_OK;
}

extern "C" int caelo_bricks_gather(caelo_ctx *ctx, const void *kpts, int kpts_f64, const int32_t *n_kpts, int F, int K,
                                   uint32_t *packed, float *patches_f32, uint8_t *trunc, const int32_t *nvox,
                                   int32_t *status, void *stream)
{
    if (!ctx || !kpts || !packed || F <= 0 || K <= 0 || (nvox && !status)) return CAELO_ERR_ARG;
    if (!ctx->bricks_tables || ctx->bricks_frames != F) return CAELO_ERR_ARG;   // no index built for this batch
    return launch_gather(ctx, kpts, kpts_f64, n_kpts, F, K, reinterpret_cast<const Table *>(ctx->bricks_tables), nvox, status,
                         packed, patches_f32, trunc, (cudaStream_t)stream);
}

extern "C" int caelo_gather_patches(caelo_ctx *ctx, const void *kpts, int kpts_f64,
                                    const int32_t *n_kpts, int F, int K, const int16_t *vox,
                                    const int64_t *vox_offsets, uint32_t *packed, float *patches_f32,
                                    uint8_t *trunc, void *stream)
{
    if (!ctx || !kpts || !vox || !vox_offsets || !packed || F <= 0 || K <= 0) return CAELO_ERR_ARG;
    int rc = caelo_bricks_build(ctx, vox, vox_offsets, F, stream);
    if (rc) return rc;
    return caelo_bricks_gather(ctx, kpts, kpts_f64, n_kpts, F, K, packed, patches_f32, trunc, nullptr, nullptr, stream);
}

// f2+a6 fused: GetPatchesList(Pts, *Voxelization(scan)) without materialising the voxel lists.
extern "C" int caelo_gather_patches_scans(caelo_ctx *ctx, const void *kpts, int kpts_f64, const int32_t *n_kpts, int F,
                                          int K, const float *pts, const int64_t *pts_offsets, uint32_t *packed,
                                          float *patches_f32, uint8_t *trunc, int32_t *nvox, int32_t *status,
                                          void *stream)
{
    if (!ctx || !kpts || !pts || !pts_offsets || !packed || !nvox || !status || F <= 0 || K <= 0) return CAELO_ERR_ARG;
    int rc = caelo_bricks_build_scans(ctx, pts, pts_offsets, F, nvox, status, stream);
    if (rc) return rc;
    return caelo_bricks_gather(ctx, kpts, kpts_f64, n_kpts, F, K, packed, patches_f32, trunc, nvox, status, stream);
}
