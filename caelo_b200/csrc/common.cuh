// common.cuh — ctx layout, scratch management and launch helpers shared by the kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/caelo.h"

struct RespondWeights {  // SphericalRingPCRespondLayer.h5, Keras layouts
    float w1[27 * 32];   // (ky,kx,ci,co)
    float b1[32];
    float w2[32 * 8];    // (ci,co)
    float b2[8];
};

struct EncoderWeightsDev {  // EncoderModel4VoxelPatch.h5 on the device
    float *k1, *b1;         // (27,8)  tap-major, (8)
    float *k2, *b2;         // (27,8,16), (16)
    float *k3, *b3;         // (27,16,32), (32)
    float *d1, *bd1;        // (2048,200), (200)
    float *d2, *bd2;        // (200,20), (20)
};

struct Scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct ProfRec {
    const char *name;
    cudaEvent_t a, b;
};

struct caelo_ctx {
    bool prof_on = false;
    long long *dbg_timeline = nullptr;  // caelo_debug_set_timeline
    std::vector<ProfRec> prof;          // one record per launch while profiling is enabled
    std::vector<cudaEvent_t> prof_pool; // recycled events
    int device = 0;
    int num_sms = 0;
    int64_t launches = 0;
    cudaError_t last_err = cudaSuccess;
    bool have_respond = false, have_encoder = false;
    float nn_margin = 1.5258789e-5f;    // nn match: proportional term of the margin, 2^-16 (see nn_margin_E in match.cu)
    RespondWeights respond_host;
    EncoderWeightsDev enc;
    float *enc_blob = nullptr;
    __half *enc_w1t_hi = nullptr, *enc_w1t_lo = nullptr;  // dense1 weights, transposed split fp16 [208][2048]
    float *enc_c12_tables = nullptr;                      // conv12: conv1 partial-sum table [9][8][8] + background table [27][16]
    // ring of pinned host staging slots for small async H2D copies (caelo_stage_acquire)
    static constexpr int kStageSlots = 8;
    void *stage_ptr[kStageSlots] = {};
    size_t stage_bytes[kStageSlots] = {};
    cudaEvent_t stage_ev[kStageSlots] = {};
    int stage_next = 0;
    // scratch regions (grown on demand, never shrunk)
    Scratch cand;      // select: candidate keys + counters
    Scratch bricks;    // patches: hash tables
    void *bricks_tables = nullptr;  // Table array of the last caelo_bricks_build* (inside `bricks`)
    int bricks_frames = 0;
    Scratch enc_ws;    // encoder activations
    Scratch pose_ws;   // ransac: hypotheses
    Scratch misc;
    Scratch scan_ws;   // projection / voxelisation: pixel owners, hash tables, compaction lists
    Scratch seed_ws;   // ransac: per-pair generator seeds
    Scratch match_ops; // nn match: split-fp16 operand tiles + padded norms
    Scratch icp_ws;    // batched ICP: grid index over PC0, nearest-neighbour indices, per-pair state
};

#define CAELO_CUDA(ctx, call)                         \
    do {                                              \
        cudaError_t e__ = (call);                     \
        if (e__ != cudaSuccess) {                     \
            (ctx)->last_err = e__;                    \
            return CAELO_ERR_CUDA;                    \
        }                                             \
    } while (0)

static inline int caelo_reserve(caelo_ctx *ctx, Scratch &s, size_t bytes)
{
    if (s.bytes >= bytes) return CAELO_OK;
    if (s.ptr) {
        // the old block may still be in use by work queued on a stream
        CAELO_CUDA(ctx, cudaDeviceSynchronize());
        CAELO_CUDA(ctx, cudaFree(s.ptr));
        s.ptr = nullptr;
        s.bytes = 0;
    }
    size_t want = bytes + bytes / 4;
    CAELO_CUDA(ctx, cudaMalloc(&s.ptr, want));
    s.bytes = want;
    return CAELO_OK;
}

// Device-memory fill by a KERNEL on the stream.  cudaMemsetAsync may be executed by a copy engine: behind a long
// host->device transfer on another stream it then waits for that transfer (measured in round 2: with the next batch's
// 84 MB upload in flight the memsets of a step stalled the compute stream by ~1 ms per step).
cudaError_t caelo_fill_async(void *ptr, int byte_value, size_t nbytes, cudaStream_t st);
// Small host -> device copy by a KERNEL that reads the pinned staging slot (caelo_stage_acquire) over PCIe: a
// cudaMemcpyAsync of a few hundred bytes on the compute stream queues on the host->device copy engine BEHIND the next
// batch's 65-85 MB upload and holds the compute stream up until that has finished.  `pinned_src` must be a slot
// returned by caelo_stage_acquire (device-accessible pinned memory), nbytes a multiple of 4.
cudaError_t caelo_stage_copy_async(void *dst, const void *pinned_src, size_t nbytes, cudaStream_t st);

#define CAELO_LAUNCH_CHECK(ctx)                       \
    do {                                              \
        (ctx)->launches++;                            \
        cudaError_t e__ = cudaGetLastError();         \
        if (e__ != cudaSuccess) {                     \
            (ctx)->last_err = e__;                    \
            return CAELO_ERR_CUDA;                    \
        }                                             \
    } while (0)


// A pinned host slot that is safe to overwrite (its previous copy has completed); the caller fills it,
// enqueues cudaMemcpyAsync from it and records *ev on the same stream.
static inline int caelo_stage_acquire(caelo_ctx *ctx, size_t bytes, void **host, cudaEvent_t *ev)
{
    const int i = ctx->stage_next;
    ctx->stage_next = (i + 1) % caelo_ctx::kStageSlots;
    if (!ctx->stage_ev[i]) CAELO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
    else CAELO_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[i]));
    if (ctx->stage_bytes[i] < bytes) {
        if (ctx->stage_ptr[i]) CAELO_CUDA(ctx, cudaFreeHost(ctx->stage_ptr[i]));
        ctx->stage_ptr[i] = nullptr;
        size_t want = bytes < 4096 ? 4096 : bytes * 2;
        CAELO_CUDA(ctx, cudaMallocHost(&ctx->stage_ptr[i], want));
        ctx->stage_bytes[i] = want;
    }
    *host = ctx->stage_ptr[i];
    *ev = ctx->stage_ev[i];
    return CAELO_OK;
}

// Per-launch device timing (caelo_profile_enable): events on the launch stream around one kernel.
struct ProfScope {
    caelo_ctx *c;
    cudaStream_t st;
    int idx = -1;
    ProfScope(caelo_ctx *ctx, const char *name, cudaStream_t stream) : c(ctx), st(stream)
    {
        if (!c->prof_on) return;
        ProfRec r;
        r.name = name;
        for (cudaEvent_t *e : {&r.a, &r.b}) {
            if (!c->prof_pool.empty()) { *e = c->prof_pool.back(); c->prof_pool.pop_back(); }
            else if (cudaEventCreate(e) != cudaSuccess) return;
        }
        cudaEventRecord(r.a, st);
        c->prof.push_back(r);
        idx = (int)c->prof.size() - 1;
    }
    ~ProfScope()
    {
        if (idx >= 0) cudaEventRecord(c->prof[idx].b, st);
    }
};
