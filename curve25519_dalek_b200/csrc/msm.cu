// msm.cu -- bucket-method (Pippenger) multiscalar multiplication on one B200.
//
// Replaces curve25519-dalek/src/backend/serial/scalar_mul/pippenger.rs:67-160 (and, being
// size-agnostic, the vartime Straus path straus.rs:159-200 the reference dispatches to below 190
// points, src/edwards.rs:1025-1029).  The reference walks 33-43 windows of 6-8 bits serially;
// here all windows run at once with c = 4..20-bit signed digits:
//
//   k_digits            one thread per scalar: signed radix-2^c digits (scalar.rs:1093-1150
//                       generalised to c > 8), histogram of bucket sizes, rank inside the bucket
//   k_scan_window       one CTA per window: exclusive scan of the bucket sizes
//   k_scatter           counting-sort scatter: point indices grouped by (window, bucket)
//   k_bucket_accumulate one thread per bucket: sum of its points with the complete unified
//                       mixed addition (curve_models.rs:411-494), 7M (affine Niels) or 8M
//   k_reduce_level      sum_k k*B_k per window by chunked running sums (pippenger.rs:146-151),
//                       log-depth across chunks
//   k_combine           total = total*2^c + window (pippenger.rs:159), compress
//
// Data layout in HBM: scalars n x 32 B; points packed Niels (96 B) or projective Niels (128 B),
// canonical 32-byte coordinates, 16-byte aligned for 128-bit loads; digit/rank entries 8 B per
// (window, scalar); sorted indices 4 B per entry; bucket sums 160 B (10 x u32 limbs x 4).
#include <algorithm>
#include <cstdio>

#include "../../include/dalek_b200.h"
#include "engine.h"

// ------------------------------------------------------------------------------------------
static inline unsigned cdiv(size_t a, unsigned b) { return (unsigned)((a + b - 1) / b); }

int ws_reserve(dalek_b200_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { ctx->last_error = std::string("cudaMalloc: ") + cudaGetErrorString(e); return DALEK_E_NOMEM; }
    b.cap = want;
    return 0;
}

int pinned_reserve(dalek_b200_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->h_pinned_cap) return 0;
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pinned = nullptr; ctx->h_pinned_cap = 0;
    cudaError_t e = cudaMallocHost(&ctx->h_pinned, bytes + 4096);
    if (e != cudaSuccess) { ctx->last_error = std::string("cudaMallocHost: ") + cudaGetErrorString(e); return DALEK_E_NOMEM; }
    ctx->h_pinned_cap = bytes + 4096;
    return 0;
}

// ------------------------------------------------------------------------------------------
// point preparation
__global__ void k_prep_compressed(const uint4 *__restrict__ in, ge_niels_packed *__restrict__ out, size_t n,
                                  int *__restrict__ bad)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a = in[2 * i], b = in[2 * i + 1];
    uint32_t s[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    fe x, y;
    uint32_t ok = ge_decompress_affine(x, y, s);
    if (!ok) { atomicOr(bad, 1); fe_0(x); fe_1(y); }   // identity placeholder keeps the kernels total
    ge_niels nl; ge_affine_to_niels(nl, x, y);
    ge_niels_packed p; ge_niels_pack(p, nl);
    uint4 *o = reinterpret_cast<uint4 *>(out + i);
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = make_uint4(p.w[4 * k], p.w[4 * k + 1], p.w[4 * k + 2], p.w[4 * k + 3]);
}

__global__ void k_prep_extended(const uint64_t *__restrict__ in, ge_pniels_packed *__restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(in + 20 * i);
    uint64_t l[20];
#pragma unroll
    for (int k = 0; k < 10; k++) { ulonglong2 v = src[k]; l[2 * k] = v.x; l[2 * k + 1] = v.y; }
    ge_p3 p;
    fe_from_limbs51(p.X, l); fe_from_limbs51(p.Y, l + 5); fe_from_limbs51(p.Z, l + 10); fe_from_limbs51(p.T, l + 15);
    ge_pniels pn; ge_p3_to_pniels(pn, p);
    ge_pniels_packed pk; ge_pniels_pack(pk, pn);
    uint4 *o = reinterpret_cast<uint4 *>(out + i);
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = make_uint4(pk.w[4 * k], pk.w[4 * k + 1], pk.w[4 * k + 2], pk.w[4 * k + 3]);
}

int msm_prepare_points(dalek_b200_ctx *ctx, const void *d_in, int point_fmt, size_t n, void *d_out, int *d_bad)
{
    if (n == 0) return 0;
    if (point_fmt == DALEK_POINTS_COMPRESSED) {
        k_prep_compressed<<<cdiv(n, 128), 128, 0, ctx->stream>>>((const uint4 *)d_in, (ge_niels_packed *)d_out, n, d_bad);
    } else {
        k_prep_extended<<<cdiv(n, 128), 128, 0, ctx->stream>>>((const uint64_t *)d_in, (ge_pniels_packed *)d_out, n);
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// window selection: minimise windows * (n + ~2.5 * buckets) (bucket adds + reduction adds)
int msm_window_count_for_bits(int c) { return 256 / c + 1; }

int msm_choose_window_bits(const dalek_b200_ctx *ctx, size_t n)
{
    if (ctx->opt_window_bits >= 4 && ctx->opt_window_bits <= 20) return (int)ctx->opt_window_bits;
    int best = 4; double best_cost = 1e300;
    for (int c = 4; c <= 20; c++) {
        double W = (double)((253 + c - 1) / c);
        double cost = W * ((double)n + 2.6 * (double)(1u << (c - 1)));
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

// ------------------------------------------------------------------------------------------
// digits + histogram.  entry = (int32 digit << 32) | rank
__global__ void k_digits(const uint4 *__restrict__ scalars, size_t n, int c, int nwin, uint32_t nbuckets,
                         uint32_t *__restrict__ counts, uint64_t *__restrict__ entries)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a = scalars[2 * i], b = scalars[2 * i + 1];
    uint32_t s[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, 0};
    const uint32_t mask = (1u << c) - 1, half = 1u << (c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < nwin; w++) {
        int o = w * c;
        uint32_t raw = 0;
        if (o < 256) {
            int wi = o >> 5, bi = o & 31;
            uint64_t two = (uint64_t)s[wi] | ((uint64_t)s[wi + 1] << 32);
            raw = (uint32_t)(two >> bi) & mask;
            if (o + c > 256) raw &= (1u << (256 - o)) - 1;
        }
        uint32_t v = raw + carry;
        int32_t d;
        if (v > half) { d = (int32_t)v - (int32_t)(1u << c); carry = 1; } else { d = (int32_t)v; carry = 0; }
        uint32_t rank = 0;
        if (d != 0) {
            uint32_t bkt = (uint32_t)(d < 0 ? -d : d) - 1;
            rank = atomicAdd(&counts[(size_t)w * nbuckets + bkt], 1u);
        }
        entries[(size_t)w * n + i] = ((uint64_t)(uint32_t)d << 32) | rank;
    }
}

// one CTA per window: exclusive scan of counts -> offsets (relative to the window's segment)
__global__ void k_scan_window(const uint32_t *__restrict__ counts, uint32_t *__restrict__ offsets, uint32_t nbuckets)
{
    __shared__ uint32_t sh[1024];
    const uint32_t *cin = counts + (size_t)blockIdx.x * nbuckets;
    uint32_t *out = offsets + (size_t)blockIdx.x * nbuckets;
    uint32_t per = (nbuckets + blockDim.x - 1) / blockDim.x;
    uint32_t lo = threadIdx.x * per, hi = min(lo + per, nbuckets);
    uint32_t sum = 0;
    for (uint32_t k = lo; k < hi; k++) sum += cin[k];
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
        uint32_t v = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = sh[threadIdx.x] - sum;
    for (uint32_t k = lo; k < hi; k++) { out[k] = run; run += cin[k]; }
}

__global__ void k_scatter(const uint64_t *__restrict__ entries, const uint32_t *__restrict__ offsets, size_t n,
                          int nwin, uint32_t nbuckets, uint32_t *__restrict__ sorted)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < nwin; w++) {
        uint64_t e = entries[(size_t)w * n + i];
        int32_t d = (int32_t)(e >> 32);
        if (d == 0) continue;
        uint32_t neg = d < 0, bkt = (uint32_t)(neg ? -d : d) - 1;
        uint32_t pos = offsets[(size_t)w * nbuckets + bkt] + (uint32_t)e;
        sorted[(size_t)w * n + pos] = (uint32_t)i | (neg << 31);
    }
}

// ------------------------------------------------------------------------------------------
// bucket accumulation: one thread per (window, bucket)
template <int KIND>
__global__ void __launch_bounds__(128)
k_bucket_accumulate(const void *__restrict__ points, const uint32_t *__restrict__ sorted,
                    const uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets, size_t n,
                    uint32_t nbuckets, uint32_t total_buckets, ge_p3_raw *__restrict__ buckets)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_buckets) return;
    uint32_t w = t / nbuckets;
    uint32_t cnt = counts[t];
    const uint32_t *idx = sorted + (size_t)w * n + offsets[t];
    ge_p3 acc; ge_p3_identity(acc);
    for (uint32_t k = 0; k < cnt; k++) {
        uint32_t e = idx[k];
        uint32_t neg = e >> 31, pi = e & 0x7fffffffu;
        if (KIND == PK_NIELS) {
            const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const ge_niels_packed *>(points) + pi);
            ge_niels_packed pk;
#pragma unroll
            for (int q = 0; q < 6; q++) { uint4 v = __ldg(src + q); pk.w[4 * q] = v.x; pk.w[4 * q + 1] = v.y; pk.w[4 * q + 2] = v.z; pk.w[4 * q + 3] = v.w; }
            ge_niels nl; ge_niels_unpack(nl, pk);
            ge_madd(acc, acc, nl, neg);
        } else {
            const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const ge_pniels_packed *>(points) + pi);
            ge_pniels_packed pk;
#pragma unroll
            for (int q = 0; q < 8; q++) { uint4 v = __ldg(src + q); pk.w[4 * q] = v.x; pk.w[4 * q + 1] = v.y; pk.w[4 * q + 2] = v.z; pk.w[4 * q + 3] = v.w; }
            ge_pniels pn; ge_pniels_unpack(pn, pk);
            ge_padd(acc, acc, pn, neg);
        }
    }
    ge_p3_raw r; ge_p3_store_raw(r, acc);
    uint4 *o = reinterpret_cast<uint4 *>(buckets + t);
#pragma unroll
    for (int q = 0; q < 10; q++) o[q] = make_uint4(r.w[4 * q], r.w[4 * q + 1], r.w[4 * q + 2], r.w[4 * q + 3]);
}

// ------------------------------------------------------------------------------------------
// weighted bucket reduction.  State at a level: items S_j and plain-sum carries V_j with
//   target = M * sum_j j*S_j + sum_j V_j          (0-based weights; M = product of earlier chunk sizes)
// One thread folds a chunk of m items:  S'_q = sum_r S_{qm+r},
//   V'_q = sum_r V_{qm+r} + M * sum_r r*S_{qm+r}   (running sums, pippenger.rs:146-151)
__device__ __forceinline__ void load_p3(ge_p3 &p, const ge_p3_raw *src)
{
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    ge_p3_raw r;
#pragma unroll
    for (int q = 0; q < 10; q++) { uint4 v = s[q]; r.w[4 * q] = v.x; r.w[4 * q + 1] = v.y; r.w[4 * q + 2] = v.z; r.w[4 * q + 3] = v.w; }
    ge_p3_load_raw(p, r);
}
__device__ __forceinline__ void store_p3(ge_p3_raw *dst, const ge_p3 &p)
{
    ge_p3_raw r; ge_p3_store_raw(r, p);
    uint4 *o = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int q = 0; q < 10; q++) o[q] = make_uint4(r.w[4 * q], r.w[4 * q + 1], r.w[4 * q + 2], r.w[4 * q + 3]);
}

__global__ void __launch_bounds__(64)
k_reduce_level(const ge_p3_raw *__restrict__ S_in, const ge_p3_raw *__restrict__ V_in, uint32_t n_in, uint32_t m,
               int log2M, uint32_t n_out, uint32_t nwin, ge_p3_raw *__restrict__ S_out, ge_p3_raw *__restrict__ V_out)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out * nwin) return;
    uint32_t w = t / n_out, q = t % n_out;
    const ge_p3_raw *S = S_in + (size_t)w * n_in + (size_t)q * m;
    uint32_t cnt = min(m, n_in - q * m);
    ge_p3 run, acc, x;
    load_p3(run, S + (cnt - 1));
    if (cnt > 1) {
        acc = run;
        for (uint32_t r = cnt - 1; r-- > 1;) {
            load_p3(x, S + r);
            ge_add(run, run, x);
            ge_add(acc, acc, run);
        }
        load_p3(x, S);
        ge_add(run, run, x);
        if (log2M > 0) ge_mul_by_pow_2(acc, acc, log2M);
    } else {
        ge_p3_identity(acc);
    }
    if (V_in) {
        const ge_p3_raw *V = V_in + (size_t)w * n_in + (size_t)q * m;
        for (uint32_t r = 0; r < cnt; r++) { load_p3(x, V + r); ge_add(acc, acc, x); }
    }
    store_p3(S_out + t, run);
    store_p3(V_out + t, acc);
}

// window accumulator = sum_j (j+1) B_j = V_top + S_top
__global__ void k_finish_windows(const ge_p3_raw *__restrict__ S_top, const ge_p3_raw *__restrict__ V_top, uint32_t nwin,
                                 ge_p3_raw *__restrict__ out)
{
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwin) return;
    ge_p3 s, v;
    load_p3(s, S_top + w);
    if (V_top) { load_p3(v, V_top + w); ge_add(s, s, v); }
    store_p3(out + w, s);
}

// Horner over windows (pippenger.rs:159) for the sum over `ranks` shards, then encode.
__global__ void k_combine(const ge_p3_raw *__restrict__ windows, int ranks, int nwin, int c, MsmResult *__restrict__ res)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    ge_p3 total, x;
    ge_p3_identity(total);
    for (int w = nwin - 1; w >= 0; w--) {
        if (w != nwin - 1) ge_mul_by_pow_2(total, total, c);
        for (int r = 0; r < ranks; r++) { load_p3(x, windows + (size_t)r * nwin + w); ge_add(total, total, x); }
    }
    uint32_t s[8];
    ge_compress(s, total);
#pragma unroll
    for (int i = 0; i < 8; i++) res->compressed[i] = s[i];
    fe_to_limbs51(res->limbs, total.X); fe_to_limbs51(res->limbs + 5, total.Y);
    fe_to_limbs51(res->limbs + 10, total.Z); fe_to_limbs51(res->limbs + 15, total.T);
    res->is_identity = ge_is_identity(total);
    res->pad = 0;
}

// ------------------------------------------------------------------------------------------
int msm_window_sums(dalek_b200_ctx *ctx, const uint32_t *d_scalars, const void *d_points, int point_kind, size_t n,
                    int c, ge_p3_raw *d_windows)
{
    const int nwin = msm_window_count_for_bits(c);
    const uint32_t nb = 1u << (c - 1);
    const size_t total_buckets = (size_t)nwin * nb;
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ws_reserve(ctx, ctx->counts, total_buckets * 4))) return rc;
    if ((rc = ws_reserve(ctx, ctx->offsets, total_buckets * 4))) return rc;
    if ((rc = ws_reserve(ctx, ctx->digits, std::max<size_t>(1, n) * nwin * 8))) return rc;
    if ((rc = ws_reserve(ctx, ctx->sorted, std::max<size_t>(1, n) * nwin * 4))) return rc;
    if ((rc = ws_reserve(ctx, ctx->buckets, total_buckets * sizeof(ge_p3_raw)))) return rc;
    uint32_t *counts = (uint32_t *)ctx->counts.p, *offsets = (uint32_t *)ctx->offsets.p;
    uint64_t *entries = (uint64_t *)ctx->digits.p;
    uint32_t *sorted = (uint32_t *)ctx->sorted.p;
    ge_p3_raw *buckets = (ge_p3_raw *)ctx->buckets.p;

    CUDA_TRY(ctx, cudaMemsetAsync(counts, 0, total_buckets * 4, st));
    if (n) {
        k_digits<<<cdiv(n, 256), 256, 0, st>>>((const uint4 *)d_scalars, n, c, nwin, nb, counts, entries);
        ctx->launches++;
    }
    k_scan_window<<<nwin, 1024, 0, st>>>(counts, offsets, nb);
    ctx->launches++;
    if (n) {
        k_scatter<<<cdiv(n, 256), 256, 0, st>>>(entries, offsets, n, nwin, nb, sorted);
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, st));
    if (point_kind == PK_NIELS)
        k_bucket_accumulate<PK_NIELS><<<cdiv(total_buckets, 128), 128, 0, st>>>(d_points, sorted, counts, offsets, n, nb, (uint32_t)total_buckets, buckets);
    else
        k_bucket_accumulate<PK_PNIELS><<<cdiv(total_buckets, 128), 128, 0, st>>>(d_points, sorted, counts, offsets, n, nb, (uint32_t)total_buckets, buckets);
    ctx->launches++;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, st));
    ctx->last_kernel_launches = 1;

    // reduction levels
    uint32_t n_in = nb;
    int log2M = 0;
    const ge_p3_raw *S_in = buckets, *V_in = nullptr;
    DevBuf *bufs[4] = {&ctx->red_a, &ctx->red_b, &ctx->red_c, &ctx->red_d};
    int flip = 0;
    bool first = true;
    while (n_in > 1) {
        uint32_t m = first ? std::min<uint32_t>(n_in, 16) : std::min<uint32_t>(n_in, 8);
        uint32_t n_out = (n_in + m - 1) / m;
        DevBuf *bs = bufs[flip], *bv = bufs[flip + 1];
        if ((rc = ws_reserve(ctx, *bs, (size_t)n_out * nwin * sizeof(ge_p3_raw)))) return rc;
        if ((rc = ws_reserve(ctx, *bv, (size_t)n_out * nwin * sizeof(ge_p3_raw)))) return rc;
        k_reduce_level<<<cdiv((size_t)n_out * nwin, 64), 64, 0, st>>>(S_in, V_in, n_in, m, log2M, n_out, nwin,
                                                                       (ge_p3_raw *)bs->p, (ge_p3_raw *)bv->p);
        ctx->launches++;
        S_in = (const ge_p3_raw *)bs->p; V_in = (const ge_p3_raw *)bv->p;
        int lg = 0; while ((1u << lg) < m) lg++;
        log2M += lg;
        n_in = n_out;
        flip = 2 - flip;
        first = false;
    }
    k_finish_windows<<<cdiv(nwin, 64), 64, 0, st>>>(S_in, V_in, nwin, d_windows);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

int msm_combine_windows(dalek_b200_ctx *ctx, const ge_p3_raw *d_windows, int ranks, int nwin, int c, MsmResult *d_result)
{
    k_combine<<<1, 32, 0, ctx->stream>>>(d_windows, ranks, nwin, c, d_result);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}
