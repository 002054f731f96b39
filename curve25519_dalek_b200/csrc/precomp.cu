// precomp.cu -- VartimePrecomputedMultiscalarMul (curve25519-dalek/src/traits.rs:290-406) for EdwardsPoint and
// RistrettoPoint (VartimeEdwardsPrecomputation, src/edwards.rs:1038-1076; VartimeRistrettoPrecomputation,
// src/ristretto.rs:1004-1049; serial backend precomputed_straus.rs:33-127).
//
// The reference precomputes width-8 NAF tables of the static points so that repeated calls skip that work.
// Here "precomputation" is what the bucket MSM can reuse between calls: the static points decoded, converted to
// packed (projective) Niels form and RESIDENT in HBM.  A call then moves only scalars (32 B per static point
// instead of 192 B) and runs two chunks onto the same buckets -- static terms, then dynamic terms -- followed by
// one reduction.  The result is the same group element as the reference's (tests compare canonical encodings).
#include <algorithm>
#include <cstring>
#include <new>

#include "../../include/dalek_b200.h"
#include "engine.h"

struct dalek_b200_precomp {
    dalek_b200_ctx *ctx;
    void *d_points;        // n packed points, device
    int kind;              // PK_NIELS / PK_PNIELS
    int ristretto;         // 1: inputs/outputs are Ristretto encodings
    size_t n;
};

static size_t in_bytes(int fmt) { return fmt == DALEK_POINTS_EXTENDED ? 160 : 32; }
static int kind_of(int fmt) { return fmt == DALEK_POINTS_COMPRESSED ? PK_NIELS : PK_PNIELS; }
static size_t packed_bytes(int kind) { return kind == PK_NIELS ? sizeof(ge_niels_packed) : sizeof(ge_pniels_packed); }

// device input in format `fmt` -> packed points at d_out
static int prepare(dalek_b200_ctx *ctx, const void *d_in, int fmt, size_t n, void *d_out, int *d_bad)
{
    if (fmt == DALEK_POINTS_RISTRETTO) return ristretto_prepare_points(ctx, d_in, n, d_out, d_bad);
    return msm_prepare_points(ctx, d_in, fmt, n, d_out, d_bad);
}

extern "C" {

int dalek_b200_precomp_new(dalek_b200_ctx *ctx, const void *static_points, int point_fmt, size_t n, dalek_b200_precomp **out)
{
    if (!ctx || !out || (n && !static_points) || n >= (1ull << 31) ||
        (point_fmt != DALEK_POINTS_COMPRESSED && point_fmt != DALEK_POINTS_EXTENDED && point_fmt != DALEK_POINTS_RISTRETTO))
        return DALEK_E_INVALID_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc;
    cudaStream_t st = ctx->stream;
    dalek_b200_precomp *pre = new (std::nothrow) dalek_b200_precomp();
    if (!pre) return DALEK_E_NOMEM;
    pre->ctx = ctx; pre->n = n; pre->kind = kind_of(point_fmt); pre->ristretto = point_fmt == DALEK_POINTS_RISTRETTO;
    pre->d_points = nullptr;
    if (cudaMalloc(&pre->d_points, std::max<size_t>(1, n) * packed_bytes(pre->kind)) != cudaSuccess) {
        ctx->last_error = "cudaMalloc failed for the static point table";
        delete pre;
        return DALEK_E_NOMEM;
    }
    auto fail = [&](int code) { cudaFree(pre->d_points); delete pre; return code; };
    if ((rc = ws_reserve(ctx, ctx->points_in, std::max<size_t>(1, n) * in_bytes(point_fmt)))) return fail(rc);
    if ((rc = ws_reserve(ctx, ctx->flags, 64))) return fail(rc);
    if ((rc = pinned_reserve(ctx, 256))) return fail(rc);
    int *h_bad = (int *)ctx->h_pinned;
    *h_bad = 0;
    if (cudaMemsetAsync(ctx->flags.p, 0, 64, st) != cudaSuccess) return fail(DALEK_E_CUDA);
    if (n && cudaMemcpyAsync(ctx->points_in.p, static_points, n * in_bytes(point_fmt), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return fail(DALEK_E_CUDA);
    if ((rc = prepare(ctx, ctx->points_in.p, point_fmt, n, pre->d_points, (int *)ctx->flags.p))) return fail(rc);
    if (cudaMemcpyAsync(h_bad, ctx->flags.p, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(DALEK_E_CUDA);
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(DALEK_E_CUDA);
    if (*h_bad) { ctx->last_error = "a static point does not decode"; return fail(DALEK_NONE); }
    *out = pre;
    return DALEK_OK;
}

size_t dalek_b200_precomp_len(const dalek_b200_precomp *pre) { return pre ? pre->n : 0; }

void dalek_b200_precomp_destroy(dalek_b200_precomp *pre)
{
    if (!pre) return;
    cudaSetDevice(pre->ctx->device);
    cudaStreamSynchronize(pre->ctx->stream);
    if (pre->d_points) cudaFree(pre->d_points);
    delete pre;
}

int dalek_b200_precomp_mixed_msm(dalek_b200_ctx *ctx, const dalek_b200_precomp *pre, const uint8_t *static_scalars, size_t n_static,
                                 const uint8_t *dynamic_scalars, const void *dynamic_points, int dynamic_fmt, size_t n_dynamic,
                                 uint8_t out_compressed[32], uint64_t out_limbs[20])
{
    if (!ctx || !pre || pre->ctx != ctx || (n_static && !static_scalars) || (n_dynamic && (!dynamic_scalars || !dynamic_points)))
        return DALEK_E_INVALID_ARG;
    if (n_static > pre->n) { ctx->last_error = "more static scalars than static points (traits.rs:317-319)"; return DALEK_E_INVALID_ARG; }
    if (dynamic_fmt != DALEK_POINTS_COMPRESSED && dynamic_fmt != DALEK_POINTS_EXTENDED && dynamic_fmt != DALEK_POINTS_RISTRETTO)
        return DALEK_E_INVALID_ARG;
    if (n_dynamic && (dynamic_fmt == DALEK_POINTS_RISTRETTO) != (pre->ristretto != 0) && dynamic_fmt != DALEK_POINTS_EXTENDED)
        return DALEK_E_INVALID_ARG;                     // Edwards and Ristretto encodings do not mix
    if (n_static + n_dynamic >= (1ull << 31)) return DALEK_E_INVALID_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CallTimer timer(ctx);
    int rc;
    cudaStream_t st = ctx->stream;
    const int dkind = kind_of(dynamic_fmt);
    const size_t din = in_bytes(dynamic_fmt);
    const int c = msm_choose_window_bits(ctx, n_static + n_dynamic);
    const int nwin = msm_window_count_for_bits(c);
    if ((rc = ws_reserve(ctx, ctx->scalars, std::max<size_t>(1, n_static + n_dynamic) * 32))) return rc;
    if ((rc = ws_reserve(ctx, ctx->points_in, std::max<size_t>(1, n_dynamic) * din))) return rc;
    if ((rc = ws_reserve(ctx, ctx->points, std::max<size_t>(1, n_dynamic) * packed_bytes(dkind)))) return rc;
    if ((rc = ws_reserve(ctx, ctx->flags, 64))) return rc;
    if ((rc = ws_reserve(ctx, ctx->misc0, (size_t)nwin * sizeof(ge_p3_raw)))) return rc;
    if ((rc = ws_reserve(ctx, ctx->result, sizeof(MsmResult) + 64))) return rc;
    if ((rc = pinned_reserve(ctx, sizeof(MsmResult) + 128))) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->flags.p, 0, 64, st));
    uint32_t *d_ss = (uint32_t *)ctx->scalars.p, *d_ds = d_ss + 8 * n_static;
    // the dynamic inputs cross PCIe on the copy stream while the static chunk is accumulated
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_copy, ctx->ev_fork, 0));
    if (n_static) CUDA_TRY(ctx, cudaMemcpyAsync(d_ss, static_scalars, n_static * 32, cudaMemcpyHostToDevice, ctx->stream_copy));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_grp[0], ctx->stream_copy));
    if (n_dynamic) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_ds, dynamic_scalars, n_dynamic * 32, cudaMemcpyHostToDevice, ctx->stream_copy));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->points_in.p, dynamic_points, n_dynamic * din, cudaMemcpyHostToDevice, ctx->stream_copy));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_grp[1], ctx->stream_copy));
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_grp[0], 0));
    if ((rc = msm_accumulate_chunk(ctx, d_ss, pre->d_points, pre->kind, n_static, c, true))) return rc;
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_grp[1], 0));
    if (n_dynamic) {
        if ((rc = prepare(ctx, ctx->points_in.p, dynamic_fmt, n_dynamic, ctx->points.p, (int *)ctx->flags.p))) return rc;
        if ((rc = msm_accumulate_chunk(ctx, d_ds, ctx->points.p, dkind, n_dynamic, c, false))) return rc;
    }
    MsmResult *d_res = (MsmResult *)ctx->result.p;
    if ((rc = msm_reduce_finish(ctx, c, (ge_p3_raw *)ctx->misc0.p, d_res))) return rc;
    uint32_t *d_enc = (uint32_t *)((char *)ctx->result.p + sizeof(MsmResult));
    if (pre->ristretto && (rc = ristretto_encode_result(ctx, d_res, d_enc))) return rc;
    MsmResult *h = (MsmResult *)ctx->h_pinned;
    int *h_bad = (int *)((char *)ctx->h_pinned + sizeof(MsmResult));
    uint8_t *h_enc = (uint8_t *)ctx->h_pinned + sizeof(MsmResult) + 64;
    CUDA_TRY(ctx, cudaMemcpyAsync(h, d_res, sizeof(MsmResult), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_bad, ctx->flags.p, 4, cudaMemcpyDeviceToHost, st));
    if (pre->ristretto) CUDA_TRY(ctx, cudaMemcpyAsync(h_enc, d_enc, 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b) == cudaSuccess) ctx->last_kernel_ms = ms;
    if (*h_bad) return DALEK_NONE;                      // optional_mixed_multiscalar_mul: a dynamic point was None
    if (out_compressed) memcpy(out_compressed, pre->ristretto ? h_enc : (const uint8_t *)h->compressed, 32);
    if (out_limbs) memcpy(out_limbs, h->limbs, 160);
    return DALEK_OK;
}

}  // extern "C"
