// fe_device_check.cu -- TEST-ONLY kernels: run the product's device field code (csrc/fe.cuh, fe64.cuh) on raw limb
// operands and return canonical encodings, so that tests/test_gpu_field.py can compare every operation with exact
// integer arithmetic (the semantics of curve25519-dalek/src/backend/serial/u64/field.rs:111-214 mul, :454-559
// pow2k/square, :368-450 to_bytes) -- including operands AT THE EXTREMES of the limb-size rules the headers state,
// which random canonical inputs never reach.  Not part of libdalek_b200.so.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../curve25519_dalek_b200/csrc/fe64.cuh"

enum { OP_FE_MUL = 0, OP_FE_SQ = 1, OP_FE64_MUL = 2, OP_FE64_SQ = 3, OP_FE64_CARRY = 4, OP_FE64_TO_FE = 5, OP_FE_SUB_MUL = 6,
       OP_FE64_FROM_FE = 7 };

__device__ __forceinline__ void load_fe(fe &f, const uint32_t *p) { for (int i = 0; i < 10; i++) f.v[i] = p[i]; }
__device__ __forceinline__ void load_fe64(fe64 &f, const long long *p) { for (int i = 0; i < 5; i++) f.v[i] = (double)p[i]; }

// out_words: canonical 32-byte encoding of the result; out_limbs (fe64 ops): the five result limbs as integers
__global__ void k_fe_check(int op, const void *a_, const void *b_, size_t n, uint32_t *out_words, long long *out_limbs)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe r;
    fe64 r64; bool have64 = false;
    if (op == OP_FE_MUL || op == OP_FE_SQ || op == OP_FE_SUB_MUL || op == OP_FE64_FROM_FE) {
        fe a, b;
        load_fe(a, (const uint32_t *)a_ + 10 * i);
        if (op == OP_FE_MUL) { load_fe(b, (const uint32_t *)b_ + 10 * i); fe_mul(r, a, b); }
        else if (op == OP_FE_SQ) fe_sq(r, a);
        else if (op == OP_FE_SUB_MUL) {        // (a - b + 2p) * (a + b): uncarried sums feeding a multiplication
            load_fe(b, (const uint32_t *)b_ + 10 * i);
            fe d, s; fe_sub(d, a, b); fe_add(s, a, b); fe_mul(r, d, s);
        } else { fe64_from_fe(r64, a); have64 = true; fe64_to_fe(r, r64); }
    } else {
        fe64 a, b;
        load_fe64(a, (const long long *)a_ + 5 * i);
        if (op == OP_FE64_MUL) { load_fe64(b, (const long long *)b_ + 5 * i); fe64_mul(r64, a, b); }
        else if (op == OP_FE64_SQ) fe64_sq(r64, a);
        else if (op == OP_FE64_CARRY) fe64_carry(r64, a);
        else r64 = a;                           // OP_FE64_TO_FE: conversion only
        have64 = true;
        if (op == OP_FE64_CARRY) {
            // the carried limbs are the observable; encode them through a second, independent route:
            // a multiplication by one (operand scale ~1 x 1)
            fe64 one; fe64_1(one); fe64 t; fe64_mul(t, r64, one); fe64_to_fe(r, t);
        } else fe64_to_fe(r, r64);
    }
    uint32_t w[8];
    fe_tobytes_words(w, r);
    for (int k = 0; k < 8; k++) out_words[8 * i + k] = w[k];
    if (out_limbs) for (int k = 0; k < 5; k++) out_limbs[5 * i + k] = have64 ? (long long)r64.v[k] : 0;
}

extern "C" int fe_device_check(int op, const void *a, const void *b, size_t n, uint32_t *out_words, long long *out_limbs)
{
    const bool is64 = !(op == OP_FE_MUL || op == OP_FE_SQ || op == OP_FE_SUB_MUL || op == OP_FE64_FROM_FE);
    const size_t in_bytes = n * (is64 ? 40 : 40);
    void *da = nullptr, *db = nullptr; uint32_t *dw = nullptr; long long *dl = nullptr;
    if (cudaMalloc(&da, in_bytes) || cudaMalloc(&db, in_bytes) || cudaMalloc(&dw, n * 32) || cudaMalloc(&dl, n * 40)) return -1;
    cudaMemcpy(da, a, in_bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b ? b : a, in_bytes, cudaMemcpyHostToDevice);
    k_fe_check<<<(unsigned)((n + 127) / 128), 128>>>(op, da, db, n, dw, dl);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out_words, dw, n * 32, cudaMemcpyDeviceToHost);
    if (out_limbs) cudaMemcpy(out_limbs, dl, n * 40, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dw); cudaFree(dl);
    return e == cudaSuccess ? 0 : -2;
}
