// microbench.cu -- measures the integer-pipe peaks the MSM roofline is quoted against
// (IMAD.WIDE.U32, IMAD, IADD3, DFMA issue rates) and the achieved rate of the field / point
// primitives built on them.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo
// Prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../curve25519_dalek_b200/csrc/ge.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_imad_wide(uint64_t *out, uint32_t a, uint32_t b)
{
    uint64_t c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    uint32_t x = a + threadIdx.x, y = b;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            c0 += (uint64_t)x * y; c1 += (uint64_t)x * y; c2 += (uint64_t)x * y; c3 += (uint64_t)x * y;
            c4 += (uint64_t)x * y; c5 += (uint64_t)x * y; c6 += (uint64_t)x * y; c7 += (uint64_t)x * y;
            // keep the compiler from folding: make operands depend on an accumulator cheaply
            asm volatile("" : "+r"(x), "+r"(y));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}

__global__ void k_imad_lo(uint32_t *out, uint32_t a, uint32_t b)
{
    uint32_t c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    uint32_t x = a + threadIdx.x, y = b;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            c0 = x * y + c0; c1 = x * y + c1; c2 = x * y + c2; c3 = x * y + c3;
            c4 = x * y + c4; c5 = x * y + c5; c6 = x * y + c6; c7 = x * y + c7;
            asm volatile("" : "+r"(x), "+r"(y));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}

__global__ void k_imad_hi(uint32_t *out, uint32_t a, uint32_t b)
{
    uint32_t c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    uint32_t x = a + threadIdx.x, y = b;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            c0 = __umulhi(x, c0) + y; c1 = __umulhi(x, c1) + y; c2 = __umulhi(x, c2) + y; c3 = __umulhi(x, c3) + y;
            c4 = __umulhi(x, c4) + y; c5 = __umulhi(x, c5) + y; c6 = __umulhi(x, c6) + y; c7 = __umulhi(x, c7) + y;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}

__global__ void k_iadd3(uint32_t *out, uint32_t a, uint32_t b)
{
    uint32_t c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    uint32_t x = a + threadIdx.x, y = b;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            asm volatile("add.u32 %0, %0, %8; add.u32 %1, %1, %9; add.u32 %2, %2, %8; add.u32 %3, %3, %9;"
                         "add.u32 %4, %4, %8; add.u32 %5, %5, %9; add.u32 %6, %6, %8; add.u32 %7, %7, %9;"
                         : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7) : "r"(x), "r"(y));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}

__global__ void k_lop_shf(uint32_t *out, uint32_t a, uint32_t b)
{
    uint32_t c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    uint32_t x = a + threadIdx.x, y = b;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            c0 = __funnelshift_r(c0, x, 7) ^ y; c1 = __funnelshift_r(c1, x, 9) ^ y; c2 = __funnelshift_r(c2, x, 11) ^ y; c3 = __funnelshift_r(c3, x, 13) ^ y;
            c4 = __funnelshift_r(c4, x, 5) ^ y; c5 = __funnelshift_r(c5, x, 3) ^ y; c6 = __funnelshift_r(c6, x, 17) ^ y; c7 = __funnelshift_r(c7, x, 19) ^ y;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}

__global__ void k_dfma(double *out, double a, double b)
{
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    double x = a + threadIdx.x * 1e-9, y = b;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            c0 = __fma_rz(x, c0, y); c1 = __fma_rz(x, c1, y); c2 = __fma_rz(x, c2, y); c3 = __fma_rz(x, c3, y);
            c4 = __fma_rz(x, c4, y); c5 = __fma_rz(x, c5, y); c6 = __fma_rz(x, c6, y); c7 = __fma_rz(x, c7, y);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

// IMAD.WIDE and IADD3 interleaved 1:1 -- do the two pipes co-issue?
__global__ void k_mix_wide_add(uint64_t *out, uint32_t a, uint32_t b)
{
    uint64_t c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3;
    uint32_t d0 = 1, d1 = 2, d2 = 3, d3 = 4;
    uint32_t x = a + threadIdx.x, y = b;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            c0 += (uint64_t)x * y; c1 += (uint64_t)x * y; c2 += (uint64_t)x * y; c3 += (uint64_t)x * y;
            asm volatile("add.u32 %0, %0, %4; add.u32 %1, %1, %5; add.u32 %2, %2, %4; add.u32 %3, %3, %5;"
                         : "+r"(d0), "+r"(d1), "+r"(d2), "+r"(d3) : "r"(x), "r"(y));
            asm volatile("" : "+r"(x), "+r"(y));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ d0 ^ d1 ^ d2 ^ d3;
}

constexpr int FE_ITERS = 2048;
__global__ void k_fe_mul(uint32_t *out, const uint32_t *in)
{
    fe a, b;
    for (int i = 0; i < 10; i++) { a.v[i] = in[i] + threadIdx.x; b.v[i] = in[10 + i] ^ (blockIdx.x & 0xff); }
    fe_carry(a, a); fe_carry(b, b);
#pragma unroll 1
    for (int i = 0; i < FE_ITERS; i++) { fe_mul(a, a, b); fe_mul(b, b, a); }
    uint32_t s = 0; for (int i = 0; i < 10; i++) s ^= a.v[i] ^ b.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fe_sq(uint32_t *out, const uint32_t *in)
{
    fe a, b;
    for (int i = 0; i < 10; i++) { a.v[i] = in[i] + threadIdx.x; b.v[i] = in[10 + i] ^ (blockIdx.x & 0xff); }
    fe_carry(a, a); fe_carry(b, b);
#pragma unroll 1
    for (int i = 0; i < FE_ITERS; i++) { fe_sq(a, a); fe_sq(b, b); }
    uint32_t s = 0; for (int i = 0; i < 10; i++) s ^= a.v[i] ^ b.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
constexpr int GE_ITERS = 512;
__global__ void k_ge_madd(uint32_t *out, const uint32_t *in)
{
    ge_p3 p; ge_p3_basepoint(p);
    ge_niels n;
    for (int i = 0; i < 10; i++) { n.ypx.v[i] = in[i] + threadIdx.x; n.ymx.v[i] = in[10 + i] ^ (blockIdx.x & 0xff); n.xy2d.v[i] = in[20 + i]; }
    fe_carry(n.ypx, n.ypx); fe_carry(n.ymx, n.ymx); fe_carry(n.xy2d, n.xy2d);
#pragma unroll 1
    for (int i = 0; i < GE_ITERS; i++) ge_madd(p, p, n, i & 1);
    uint32_t s = 0; for (int i = 0; i < 10; i++) s ^= p.X.v[i] ^ p.Y.v[i] ^ p.Z.v[i] ^ p.T.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ge_dbl(uint32_t *out, const uint32_t *in)
{
    ge_p3 p; ge_p3_basepoint(p);
    p.X.v[0] ^= in[0] & 1; // opaque
#pragma unroll 1
    for (int i = 0; i < GE_ITERS; i++) ge_dbl(p, p);
    uint32_t s = 0; for (int i = 0; i < 10; i++) s ^= p.X.v[i] ^ p.Y.v[i] ^ p.Z.v[i] ^ p.T.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps = 5)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    void *out; CK(cudaMalloc(&out, (size_t)sms * 64 * 1024 * 8));
    uint32_t hin[64]; for (int i = 0; i < 64; i++) hin[i] = 0x1234567u * (i + 1) & 0x3ffffff;
    uint32_t *in; CK(cudaMalloc(&in, sizeof hin)); CK(cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
    const int threads = 256;
    for (int bps : {4, 8}) {
        int blocks = sms * bps;
        double n = (double)blocks * threads;
        float ms;
        ms = time_ms([&] { k_imad_wide<<<blocks, threads>>>((uint64_t *)out, 3, 5); });
        printf(", \"imad_wide_per_clk_sm_bps%d\": %.2f", bps, n * ITERS * 32 / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
        printf(", \"imad_wide_Gops_bps%d\": %.1f", bps, n * ITERS * 32 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_imad_lo<<<blocks, threads>>>((uint32_t *)out, 3, 5); });
        printf(", \"imad_lo_Gops_bps%d\": %.1f", bps, n * ITERS * 32 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_imad_hi<<<blocks, threads>>>((uint32_t *)out, 3, 5); });
        printf(", \"imad_hi_Gops_bps%d\": %.1f", bps, n * ITERS * 32 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_iadd3<<<blocks, threads>>>((uint32_t *)out, 3, 5); });
        printf(", \"iadd_Gops_bps%d\": %.1f", bps, n * ITERS * 32 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_lop_shf<<<blocks, threads>>>((uint32_t *)out, 3, 5); });
        printf(", \"shf_lop_pairs_Gops_bps%d\": %.1f", bps, n * ITERS * 32 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_dfma<<<blocks, threads>>>((double *)out, 1.0000001, 0.5); });
        printf(", \"dfma_Gops_bps%d\": %.1f", bps, n * ITERS * 32 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_mix_wide_add<<<blocks, threads>>>((uint64_t *)out, 3, 5); });
        printf(", \"mix_wide_plus_add_Gops_each_bps%d\": %.1f", bps, n * ITERS * 32 / (ms * 1e-3) / 1e9);
    }
    for (int threads2 : {128, 256}) for (int bps : {1, 2, 4}) {
        int blocks = sms * bps; double n = (double)blocks * threads2; float ms;
        ms = time_ms([&] { k_fe_mul<<<blocks, threads2>>>((uint32_t *)out, in); });
        printf(", \"fe_mul_G_t%d_b%d\": %.2f", threads2, bps, n * FE_ITERS * 2 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_fe_sq<<<blocks, threads2>>>((uint32_t *)out, in); });
        printf(", \"fe_sq_G_t%d_b%d\": %.2f", threads2, bps, n * FE_ITERS * 2 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_ge_madd<<<blocks, threads2>>>((uint32_t *)out, in); });
        printf(", \"ge_madd_G_t%d_b%d\": %.3f", threads2, bps, n * GE_ITERS / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_ge_dbl<<<blocks, threads2>>>((uint32_t *)out, in); });
        printf(", \"ge_dbl_G_t%d_b%d\": %.3f", threads2, bps, n * GE_ITERS / (ms * 1e-3) / 1e9);
    }
    printf("}\n");
    return 0;
}
