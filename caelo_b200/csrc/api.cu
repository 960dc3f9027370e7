// api.cu — ctx lifetime, weights and error reporting for libcaelo_b200.so (include/caelo.h).
#include <stdio.h>

#include <new>

#include "common.cuh"

int caelo_encoder_init(caelo_ctx *ctx);
int caelo_encoder_prepare(caelo_ctx *ctx);
int caelo_match_init(caelo_ctx *ctx);
int caelo_pose_init(caelo_ctx *ctx);
int caelo_select_init(caelo_ctx *ctx);

namespace {
__global__ void __launch_bounds__(256) fill_kernel(unsigned char *p, unsigned char v, size_t head, size_t n16, size_t tail)
{
    // [head bytes][n16 aligned 16-byte words][tail bytes]
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    const unsigned w = v * 0x01010101u;
    uint4 *q = reinterpret_cast<uint4 *>(p + head);
    for (size_t k = i; k < n16; k += stride) q[k] = make_uint4(w, w, w, w);
    if (i < head) p[i] = v;
    if (i < tail) p[head + n16 * 16 + i] = v;
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256) stage_copy_kernel(unsigned *dst, const unsigned *src, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace

cudaError_t caelo_stage_copy_async(void *dst, const void *pinned_src, size_t nbytes, cudaStream_t st)
{
    if (!nbytes) return cudaSuccess;
    void *dsrc = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&dsrc, const_cast<void *>(pinned_src), 0);
    if (e != cudaSuccess) return e;
    const size_t n = (nbytes + 3) / 4;
    size_t blocks = (n + 255) / 256;
    if (blocks > 64) blocks = 64;
    stage_copy_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<unsigned *>(dst), static_cast<const unsigned *>(dsrc), n);
    return cudaGetLastError();
}

cudaError_t caelo_fill_async(void *ptr, int byte_value, size_t nbytes, cudaStream_t st)
{
    if (!nbytes) return cudaSuccess;
    unsigned char *p = static_cast<unsigned char *>(ptr);
    size_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > nbytes) head = nbytes;
    const size_t n16 = (nbytes - head) / 16, tail = nbytes - head - n16 * 16;
    size_t blocks = (n16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, (unsigned char)byte_value, head, n16, tail);
    return cudaGetLastError();
}

extern "C" int caelo_version(void) { return 100; }

extern "C" const char *caelo_error_string(int code)
{
    switch (code) {
    case CAELO_OK: return "ok";
    case CAELO_ERR_CUDA: return "CUDA runtime error (see caelo_last_cuda_error)";
    case CAELO_ERR_ARG: return "invalid argument";
    case CAELO_ERR_NO_WEIGHTS: return "network weights not set on this ctx";
    case CAELO_ERR_NONBINARY_PATCH: return "encoder input contains a value other than 0 or 1";
    case CAELO_ERR_TOO_FEW_VOXELS: return "a voxel list has fewer than 496 entries (n_neighbors <= n_samples_fit)";
    case CAELO_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
    case CAELO_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
    }
}

extern "C" const char *caelo_last_cuda_error(const caelo_ctx *ctx)
{
    return ctx ? cudaGetErrorString(ctx->last_err) : "null ctx";
}

extern "C" int caelo_create(int device_id, caelo_ctx **out)
{
    if (!out) return CAELO_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_id < 0 || device_id >= n)
        return CAELO_ERR_NO_DEVICE;
    if (cudaSetDevice(device_id) != cudaSuccess) return CAELO_ERR_NO_DEVICE;
    caelo_ctx *ctx = new (std::nothrow) caelo_ctx();
    if (!ctx) return CAELO_ERR_ARG;
    ctx->device = device_id;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) { delete ctx; return CAELO_ERR_NO_DEVICE; }
    ctx->num_sms = prop.multiProcessorCount;
    memset(&ctx->enc, 0, sizeof(ctx->enc));
    int rc = caelo_encoder_init(ctx);
    if (!rc) rc = caelo_match_init(ctx);
    if (!rc) rc = caelo_pose_init(ctx);
    if (!rc) rc = caelo_select_init(ctx);
    if (rc) { delete ctx; return rc; }
    *out = ctx;
    return CAELO_OK;
}

extern "C" int caelo_destroy(caelo_ctx *ctx)
{
    if (!ctx) return CAELO_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    Scratch *all[] = {&ctx->cand, &ctx->bricks, &ctx->enc_ws, &ctx->pose_ws, &ctx->misc, &ctx->scan_ws, &ctx->seed_ws, &ctx->match_ops, &ctx->icp_ws};
    for (Scratch *s : all)
        if (s->ptr) cudaFree(s->ptr);
    for (int i = 0; i < caelo_ctx::kStageSlots; ++i) {
        if (ctx->stage_ptr[i]) cudaFreeHost(ctx->stage_ptr[i]);
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    }
    if (ctx->enc_blob) cudaFree(ctx->enc_blob);
    if (ctx->enc_w1t_hi) cudaFree(ctx->enc_w1t_hi);
    if (ctx->enc_c12_tables) cudaFree(ctx->enc_c12_tables);
    delete ctx;
    return CAELO_OK;
}

extern "C" int caelo_profile_enable(caelo_ctx *ctx, int on)
{
    if (!ctx) return CAELO_ERR_ARG;
    ctx->prof_on = on != 0;
    return CAELO_OK;
}

// Writes "name count total_ms\n" lines (aggregated per kernel name, launch order of first use)
// into buf, clears the records.  Synchronises the device.
extern "C" int caelo_profile_fetch(caelo_ctx *ctx, char *buf, int buflen)
{
    if (!ctx || !buf || buflen <= 0) return CAELO_ERR_ARG;
    CAELO_CUDA(ctx, cudaDeviceSynchronize());
    struct Agg { const char *name; int n; double ms; };
    std::vector<Agg> agg;
    for (ProfRec &r : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        size_t i = 0;
        for (; i < agg.size(); ++i)
            if (!strcmp(agg[i].name, r.name)) break;
        if (i == agg.size()) agg.push_back({r.name, 0, 0.0});
        agg[i].n++;
        agg[i].ms += ms;
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof.clear();
    int pos = 0;
    buf[0] = 0;
    for (Agg &a : agg) {
        int w = snprintf(buf + pos, buflen - pos, "%s %d %.6f\n", a.name, a.n, a.ms);
        if (w < 0 || w >= buflen - pos) break;
        pos += w;
    }
    return CAELO_OK;
}

extern "C" int caelo_debug_set_timeline(caelo_ctx *ctx, long long *buf)
{
    if (!ctx) return CAELO_ERR_ARG;
    ctx->dbg_timeline = buf;
    return CAELO_OK;
}

extern "C" int caelo_num_sms(const caelo_ctx *ctx) { return ctx ? ctx->num_sms : 0; }
extern "C" int64_t caelo_launch_count(const caelo_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int caelo_set_respond_weights(caelo_ctx *ctx, const float *w1, const float *b1,
                                         const float *w2, const float *b2)
{
    if (!ctx || !w1 || !b1 || !w2 || !b2) return CAELO_ERR_ARG;
    memcpy(ctx->respond_host.w1, w1, sizeof(ctx->respond_host.w1));
    memcpy(ctx->respond_host.b1, b1, sizeof(ctx->respond_host.b1));
    memcpy(ctx->respond_host.w2, w2, sizeof(ctx->respond_host.w2));
    memcpy(ctx->respond_host.b2, b2, sizeof(ctx->respond_host.b2));
    ctx->have_respond = true;
    return CAELO_OK;
}

extern "C" int caelo_set_encoder_weights(caelo_ctx *ctx, const float *k1, const float *b1,
                                         const float *k2, const float *b2, const float *k3,
                                         const float *b3, const float *d1, const float *bd1,
                                         const float *d2, const float *bd2)
{
    if (!ctx || !k1 || !b1 || !k2 || !b2 || !k3 || !b3 || !d1 || !bd1 || !d2 || !bd2) return CAELO_ERR_ARG;
    const size_t n[10] = {27 * 8, 8, 27 * 8 * 16, 16, 27 * 16 * 32, 32, 2048 * 200, 200, 200 * 20, 20};
    const float *src[10] = {k1, b1, k2, b2, k3, b3, d1, bd1, d2, bd2};
    size_t total = 0, off[10];
    for (int i = 0; i < 10; ++i) { off[i] = total; total += (n[i] + 63) / 64 * 64; }
    CAELO_CUDA(ctx, cudaSetDevice(ctx->device));
    // the weights are context state: kernels still queued on any stream read the old ones — let them finish first
    if (ctx->have_encoder) CAELO_CUDA(ctx, cudaDeviceSynchronize());
    if (!ctx->enc_blob) CAELO_CUDA(ctx, cudaMalloc(&ctx->enc_blob, total * 4));
    for (int i = 0; i < 10; ++i)
        CAELO_CUDA(ctx, cudaMemcpy(ctx->enc_blob + off[i], src[i], n[i] * 4, cudaMemcpyHostToDevice));
    float *b = ctx->enc_blob;
    ctx->enc.k1 = b + off[0]; ctx->enc.b1 = b + off[1]; ctx->enc.k2 = b + off[2]; ctx->enc.b2 = b + off[3];
    ctx->enc.k3 = b + off[4]; ctx->enc.b3 = b + off[5]; ctx->enc.d1 = b + off[6]; ctx->enc.bd1 = b + off[7];
    ctx->enc.d2 = b + off[8]; ctx->enc.bd2 = b + off[9];
    int rc = caelo_encoder_prepare(ctx);
    if (rc) return rc;
    ctx->have_encoder = true;
    return CAELO_OK;
}
