// umma_debug.cu — bring-up/test hook for the tcgen05 path: one CTA computes D = A * B^T with
// operands staged in the canonical K-major no-swizzle layout under caller-chosen SBO/LBO and
// dumps the raw TMEM accumulator (128 lanes x 64 columns), so that tests can pin the smem
// descriptor semantics and the TMEM data-path layout the encoder kernels rely on.
#include "common.cuh"
#include "umma.cuh"

namespace {

struct DbgArgs {
    const __half *A, *B;  // [M,K], [N,K] row-major
    int M, N, K;
    int sbo_a, lbo_a, sbo_b, lbo_b;  // bytes
    int d_lane_off;                  // accumulator lane offset (0 or 16 for interleaved M=64 tiles)
    float *dump;                     // [128,64]
};

__global__ void __launch_bounds__(128) umma_debug_kernel(const DbgArgs a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char *sa = sm;
    unsigned char *sb = sm + 65536;
    for (int i = tid; i < 131072 / 16; i += 128) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) umma::tmem_alloc(&tmem_base, 64);
    if (tid == 0) {
        umma::mbar_init(&mbar, 1);
        umma::fence_mbar_init();
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t taddr = tmem_base;
    for (int e = tid; e < a.M * a.K; e += 128) {
        int r = e / a.K, k = e % a.K;
        *reinterpret_cast<__half *>(sa + (r / 8) * a.sbo_a + (k / 8) * a.lbo_a + (r % 8) * 16 + (k % 8) * 2) = a.A[e];
    }
    for (int e = tid; e < a.N * a.K; e += 128) {
        int r = e / a.K, k = e % a.K;
        *reinterpret_cast<__half *>(sb + (r / 8) * a.sbo_b + (k / 8) * a.lbo_b + (r % 8) * 16 + (k % 8) * 2) = a.B[e];
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        const uint32_t idesc = umma::idesc_f16_f32(a.M, a.N);
        for (int ks = 0; ks < a.K / 16; ++ks) {
            uint64_t da = umma::smem_desc(umma::smem_u32(sa) + 2 * ks * a.lbo_a, a.lbo_a, a.sbo_a);
            uint64_t db = umma::smem_desc(umma::smem_u32(sb) + 2 * ks * a.lbo_b, a.lbo_b, a.sbo_b);
            umma::mma_f16(taddr + ((uint32_t)a.d_lane_off << 16), da, db, idesc, ks > 0 ? 1u : 0u);
        }
        umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, 0);
    umma::fence_after_thread_sync();
    uint32_t v[32];
    for (int half = 0; half < 2; ++half) {
        umma::tmem_ld_x32(taddr + ((uint32_t)(warp * 32) << 16) + half * 32, v);
        umma::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) a.dump[(warp * 32 + (tid & 31)) * 64 + half * 32 + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(taddr, 64);
}

}  // namespace

extern "C" int caelo_debug_umma(caelo_ctx *ctx, const void *A, const void *B, int M, int N, int K, int sbo_a,
                                int lbo_a, int sbo_b, int lbo_b, int d_lane_off, float *dump, void *stream)
{
    if (!ctx || !A || !B || !dump) return CAELO_ERR_ARG;
    if ((M != 64 && M != 128) || N < 8 || N > 64 || N % 8 || K < 16 || K > 256 || K % 16) return CAELO_ERR_ARG;
    if ((M / 8 - 1) * sbo_a + (K / 8 - 1) * lbo_a + 128 > 65536) return CAELO_ERR_ARG;
    if ((N / 8 - 1) * sbo_b + (K / 8 - 1) * lbo_b + 128 > 65536) return CAELO_ERR_ARG;
    static bool attr = false;
    if (!attr) {
        CAELO_CUDA(ctx, cudaFuncSetAttribute(umma_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        attr = true;
    }
    DbgArgs a;
    a.A = (const __half *)A; a.B = (const __half *)B; a.M = M; a.N = N; a.K = K;
    a.sbo_a = sbo_a; a.lbo_a = lbo_a; a.sbo_b = sbo_b; a.lbo_b = lbo_b; a.d_lane_off = d_lane_off; a.dump = dump;
    umma_debug_kernel<<<1, 128, 131072, (cudaStream_t)stream>>>(a);
    CAELO_LAUNCH_CHECK(ctx);
    return CAELO_OK;
}
