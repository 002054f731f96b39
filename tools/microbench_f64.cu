// microbench_f64.cu -- prototype of GF(2^255-19) multiplication on the FP64 pipe (DFMA, 64/clk/SM on
// B200, twice the IMAD.WIDE rate): 5 x 51-bit limbs held as doubles, products split into
// (floor(p / 2^51), p mod 2^51) with two round-mode FMAs, hi parts accumulated as raw IEEE bit
// patterns in 64-bit integers, lo parts accumulated exactly in double.  Checks itself against the
// integer implementation (fe.cuh) and reports muls/s for both.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../curve25519_dalek_b200/csrc/ge.cuh"

struct fe64 { double v[5]; };      // integer-valued limbs, 0 <= v < 2^51 (+ small)

#define E52 0x4330000000000000LL   // bits of 2^52

__device__ __forceinline__ void fe64_from_fe(fe64 &h, const fe &f)
{
    uint32_t w[8]; fe_tobytes_words(w, f); fe t; fe_frombytes_words(t, w);
#pragma unroll
    for (int i = 0; i < 5; i++) h.v[i] = (double)((uint64_t)t.v[2 * i] | ((uint64_t)t.v[2 * i + 1] << 26));
}
__device__ __forceinline__ void fe_from_fe64(fe &f, const fe64 &h)
{
    uint64_t l[5];
#pragma unroll
    for (int i = 0; i < 5; i++) l[i] = (uint64_t)h.v[i];
    fe_from_limbs51(f, l);
}

// shared tail: nine signed int64 columns -> wrap (x19), carry, back to doubles
__device__ __forceinline__ void fe64_finish(fe64 &h, long long V[9])
{
    long long R[5];
#pragma unroll
    for (int k = 0; k < 4; k++) R[k] = V[k] + 19 * V[k + 5];
    R[4] = V[4];
    long long c;
    c = R[0] >> 51; R[1] += c; R[0] &= 0x7ffffffffffffLL;
    c = R[1] >> 51; R[2] += c; R[1] &= 0x7ffffffffffffLL;
    c = R[2] >> 51; R[3] += c; R[2] &= 0x7ffffffffffffLL;
    c = R[3] >> 51; R[4] += c; R[3] &= 0x7ffffffffffffLL;
    c = R[4] >> 51; R[0] += 19 * c; R[4] &= 0x7ffffffffffffLL;
    c = R[0] >> 51; R[1] += c; R[0] &= 0x7ffffffffffffLL;
#pragma unroll
    for (int k = 0; k < 5; k++) h.v[k] = __longlong_as_double(R[k] | E52) - 4503599627370496.0;
}

// v1: hi and lo both accumulated as raw bits (int64 adds on the ALU pipe)
__device__ __forceinline__ void fe64_mul_v1(fe64 &h, const fe64 &a, const fe64 &b)
{
    const double M1 = 6755399441055744.0;                  // 1.5 * 2^52
    const double K = 6755399441055744.0 * 2251799813685248.0 + 4503599627370496.0;   // M1*2^51 + 2^52
    double bs[5];
#pragma unroll
    for (int j = 0; j < 5; j++) bs[j] = b.v[j] * (1.0 / 2251799813685248.0);
    long long H[10], L[9];
#pragma unroll
    for (int k = 0; k < 10; k++) H[k] = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) L[k] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            double t = __fma_rz(a.v[i], bs[j], M1);                 // M1 + floor(p / 2^51)
            double u = __fma_rn(t, -2251799813685248.0, K);         // 2^52 - floor * 2^51
            double lo = __fma_rn(a.v[i], b.v[j], u);                // 2^52 + (p mod 2^51)
            H[i + j + 1] += __double_as_longlong(t);
            L[i + j] += __double_as_longlong(lo);
        }
    long long V[9];
    const long long EH = 0x4330000000000000LL + (1LL << 51);     // bits(M1) = bits(2^52) + 2^51
#pragma unroll
    for (int k = 0; k < 9; k++) {
        int nl = (k < 5 ? k + 1 : 9 - k), nh = (k == 0 ? 0 : (k - 1 < 5 ? k : 10 - k));
        V[k] = (L[k] - nl * E52) + (H[k] - nh * EH);
    }
    // column 9 only has the hi of (4,4): fold it with weight 19 into column 4
    V[4] += 19 * (H[9] - EH);
    fe64_finish(h, V);
}

// v2: lo parts un-offset and summed exactly in double (<= 4 terms per accumulator), hi as raw bits
__device__ __forceinline__ void fe64_mul_v2(fe64 &h, const fe64 &a, const fe64 &b)
{
    const double M1 = 6755399441055744.0;
    const double K0 = 6755399441055744.0 * 2251799813685248.0;      // M1 * 2^51
    double bs[5];
#pragma unroll
    for (int j = 0; j < 5; j++) bs[j] = b.v[j] * (1.0 / 2251799813685248.0);
    long long H[10];
    double L[9], L4b = 0.0;
#pragma unroll
    for (int k = 0; k < 10; k++) H[k] = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) L[k] = 0.0;
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            double t = __fma_rz(a.v[i], bs[j], M1);
            double u = __fma_rn(t, -2251799813685248.0, K0);        // -floor * 2^51
            double lo = __fma_rn(a.v[i], b.v[j], u);                // p mod 2^51, in [0, 2^51)
            H[i + j + 1] += __double_as_longlong(t);
            if (i + j == 4 && i == 4) L4b += lo; else L[i + j] += lo;   // column 4 has 5 terms: keep sums < 2^53
        }
    long long V[9];
    const long long EH = 0x4330000000000000LL + (1LL << 51);
#pragma unroll
    for (int k = 0; k < 9; k++) {
        int nh = (k == 0 ? 0 : (k - 1 < 5 ? k : 10 - k));
        V[k] = __double2ll_rz(L[k]) + (H[k] - nh * EH);
    }
    V[4] += __double2ll_rz(L4b) + 19 * (H[9] - EH);
    fe64_finish(h, V);
}

// squaring v2-style: 15 products, cross terms doubled through a pre-doubled operand
__device__ __forceinline__ void fe64_sq_v2(fe64 &h, const fe64 &a)
{
    const double M1 = 4503599627370496.0;                                  // 2^52: operands are non-negative, floor < 2^52
    const double K0 = 4503599627370496.0 * 2251799813685248.0;
    double as[5], a2[5];
#pragma unroll
    for (int j = 0; j < 5; j++) { as[j] = a.v[j] * (1.0 / 2251799813685248.0); a2[j] = a.v[j] + a.v[j]; }
    long long H[10];
    double L[9];
#pragma unroll
    for (int k = 0; k < 10; k++) H[k] = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) L[k] = 0.0;
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = i; j < 5; j++) {
            double x = (i < j) ? a2[i] : a.v[i];                    // 2 a_i a_j for cross terms: p < 2^103
            double t = __fma_rz(x, as[j], M1);                      // floor(p / 2^51) < 2^52: needs M1 range... see note
            double u = __fma_rn(t, -2251799813685248.0, K0);
            double lo = __fma_rn(x, a.v[j], u);
            H[i + j + 1] += __double_as_longlong(t);
            L[i + j] += lo;
        }
    long long V[9];
    const long long EH = 0x4330000000000000LL;
    const int nh_tab[10] = {0, 1, 1, 2, 2, 3, 2, 2, 1, 1};
#pragma unroll
    for (int k = 0; k < 9; k++) V[k] = __double2ll_rz(L[k]) + (H[k] - nh_tab[k] * EH);
    V[4] += 19 * (H[9] - EH);
    fe64_finish(h, V);
}

constexpr int ITERS = 1024;

// correctness: one warp, FP64 result vs integer result
template <int VER>
__global__ void k_check(uint64_t *out, const uint32_t *in)
{
    fe a, b;
    for (int i = 0; i < 10; i++) { a.v[i] = in[i] + threadIdx.x; b.v[i] = in[10 + i] ^ 0x5a; }
    fe_carry(a, a); fe_carry(b, b);
    fe64 x, y; fe64_from_fe(x, a); fe64_from_fe(y, b);
    for (int i = 0; i < 64; i++) {
        if (VER == 1) { fe64_mul_v1(x, x, y); fe64_mul_v1(y, y, x); fe_mul(a, a, b); fe_mul(b, b, a); }
        else if (VER == 2) { fe64_mul_v2(x, x, y); fe64_mul_v2(y, y, x); fe_mul(a, a, b); fe_mul(b, b, a); }
        else { fe64_sq_v2(x, x); fe64_sq_v2(y, y); fe_sq(a, a); fe_sq(b, b); }
    }
    fe xa, yb; fe_from_fe64(xa, x); fe_from_fe64(yb, y);
    out[threadIdx.x] = (uint64_t)(fe_eq(xa, a) & fe_eq(yb, b));
}

// MODE 0: integer mul, 1: f64 mul v1, 2: f64 mul v2, 3: f64 sq, 4: integer sq,
// 5: mixed -- even warps integer mul, odd warps f64 mul v1 (both pipes busy at once)
template <int MODE>
__global__ void __launch_bounds__(128) k_rate(uint64_t *out, const uint32_t *in)
{
    fe a, b;
    for (int i = 0; i < 10; i++) { a.v[i] = in[i] + threadIdx.x; b.v[i] = in[10 + i] ^ (blockIdx.x & 0xff); }
    fe_carry(a, a); fe_carry(b, b);
    uint64_t s = 0;
    const bool use_int = MODE == 0 || MODE == 4 || (MODE == 5 && ((threadIdx.x >> 5) & 1) == 0);
    if (use_int) {
#pragma unroll 1
        for (int i = 0; i < ITERS; i++) { if (MODE == 4) { fe_sq(a, a); fe_sq(b, b); } else { fe_mul(a, a, b); fe_mul(b, b, a); } }
        for (int i = 0; i < 10; i++) s ^= a.v[i] ^ b.v[i];
    } else {
        fe64 x, y; fe64_from_fe(x, a); fe64_from_fe(y, b);
#pragma unroll 1
        for (int i = 0; i < ITERS; i++) {
            if (MODE == 1 || MODE == 5) { fe64_mul_v1(x, x, y); fe64_mul_v1(y, y, x); }
            else if (MODE == 2) { fe64_mul_v2(x, x, y); fe64_mul_v2(y, y, x); }
            else { fe64_sq_v2(x, x); fe64_sq_v2(y, y); }
        }
        for (int i = 0; i < 5; i++) s ^= (uint64_t)__double_as_longlong(x.v[i]) ^ (uint64_t)__double_as_longlong(y.v[i]);
    }
    out[64 + blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 6 family: squarings on both pipes at once.  Of every 4 warps, IW run the integer squaring (IMAD.WIDE on
// the FMA pipe) for ITERS * PCT / 100 iterations and the others the FP64 squaring for ITERS iterations.
template <int IW, int PCT>
__global__ void __launch_bounds__(128) k_mix_sq(uint64_t *out, const uint32_t *in)
{
    fe a, b;
    for (int i = 0; i < 10; i++) { a.v[i] = in[i] + threadIdx.x; b.v[i] = in[10 + i] ^ (blockIdx.x & 0xff); }
    fe_carry(a, a); fe_carry(b, b);
    uint64_t s = 0;
    if ((int)((threadIdx.x >> 5) & 3) < IW) {
#pragma unroll 1
        for (int i = 0; i < ITERS * PCT / 100; i++) { fe_sq(a, a); fe_sq(b, b); }
        for (int i = 0; i < 10; i++) s ^= a.v[i] ^ b.v[i];
    } else {
        fe64 x, y; fe64_from_fe(x, a); fe64_from_fe(y, b);
#pragma unroll 1
        for (int i = 0; i < ITERS; i++) { fe64_sq_v2(x, x); fe64_sq_v2(y, y); }
        for (int i = 0; i < 5; i++) s ^= (uint64_t)__double_as_longlong(x.v[i]) ^ (uint64_t)__double_as_longlong(y.v[i]);
    }
    out[64 + blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> static float time_ms(F launch, int reps = 5)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    uint64_t *out; cudaMalloc(&out, (size_t)(64 + sms * 16 * 128) * 8);
    uint32_t hin[64]; for (int i = 0; i < 64; i++) hin[i] = 0x1234567u * (i + 1) & 0x3ffffff;
    uint32_t *in; cudaMalloc(&in, sizeof hin); cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    uint64_t ok[32];
    printf("{\"gpu\": \"%s\"", prop.name);
    k_check<1><<<1, 32>>>(out, in); cudaMemcpy(ok, out, sizeof ok, cudaMemcpyDeviceToHost); printf(", \"v1_ok\": %d", (int)(ok[0] & ok[7] & ok[31]));
    k_check<2><<<1, 32>>>(out, in); cudaMemcpy(ok, out, sizeof ok, cudaMemcpyDeviceToHost); printf(", \"v2_ok\": %d", (int)(ok[0] & ok[7] & ok[31]));
    k_check<3><<<1, 32>>>(out, in); cudaMemcpy(ok, out, sizeof ok, cudaMemcpyDeviceToHost); printf(", \"sq_ok\": %d", (int)(ok[0] & ok[7] & ok[31]));
    for (int bps : {2, 3, 4, 8, 16}) {
        int blocks = sms * bps, threads = 128; double n = (double)blocks * threads * ITERS * 2; float ms;
        ms = time_ms([&] { k_rate<0><<<blocks, threads>>>(out, in); }); printf(", \"int_mul_G_b%d\": %.1f", bps, n / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_rate<1><<<blocks, threads>>>(out, in); }); printf(", \"f64_mul_v1_G_b%d\": %.1f", bps, n / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_rate<2><<<blocks, threads>>>(out, in); }); printf(", \"f64_mul_v2_G_b%d\": %.1f", bps, n / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_rate<4><<<blocks, threads>>>(out, in); }); printf(", \"int_sq_G_b%d\": %.1f", bps, n / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_rate<3><<<blocks, threads>>>(out, in); }); printf(", \"f64_sq_G_b%d\": %.1f", bps, n / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_rate<5><<<blocks, threads>>>(out, in); }); printf(", \"mixed_int_f64_mul_G_b%d\": %.1f", bps, n / (ms * 1e-3) / 1e9);
    }
    for (int bps : {4, 8}) {
        int blocks = sms * bps, threads = 128; float ms;
#define MIX(IW, PCT) ms = time_ms([&] { k_mix_sq<IW, PCT><<<blocks, threads>>>(out, in); }); \
        printf(", \"mix_sq_iw%d_pct%d_G_b%d\": %.1f", IW, PCT, bps, (double)blocks * 32 * 2 * (IW * (ITERS * PCT / 100) + (4 - IW) * ITERS) / (ms * 1e-3) / 1e9);
        MIX(1, 50) MIX(1, 100) MIX(1, 150) MIX(2, 50) MIX(2, 75) MIX(2, 100) MIX(3, 50) MIX(3, 100)
#undef MIX
    }
    printf("}\n");
    return 0;
}
