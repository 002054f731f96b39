// Host build of the device field/point headers with limb-bound assertions enabled, exported with a
// tiny C ABI so tests/test_fe_host.py can compare every operation with the CPU oracle.
// TEST INFRASTRUCTURE: this is not a CPU fallback of the product (the product library refuses to
// run without a GPU); it only checks that fe.cuh / ge.cuh never exceed their limb bounds.
#define FE_CHECK_BOUNDS 1
#include "../../curve25519_dalek_b200/csrc/ge.cuh"
#include <string.h>

static void load(fe &f, const uint8_t *b) { uint32_t w[8]; memcpy(w, b, 32); fe_frombytes_words(f, w); }
static void store(uint8_t *b, const fe &f) { uint32_t w[8]; fe_tobytes_words(w, f); memcpy(b, w, 32); }

extern "C" {
void h_fe_mul(uint8_t *o, const uint8_t *a, const uint8_t *b) { fe x, y, z; load(x, a); load(y, b); fe_mul(z, x, y); store(o, z); }
void h_fe_sq(uint8_t *o, const uint8_t *a) { fe x, z; load(x, a); fe_sq(z, x); store(o, z); }
void h_fe_add(uint8_t *o, const uint8_t *a, const uint8_t *b) { fe x, y, z; load(x, a); load(y, b); fe_add(z, x, y); store(o, z); }
void h_fe_sub(uint8_t *o, const uint8_t *a, const uint8_t *b) { fe x, y, z; load(x, a); load(y, b); fe_sub(z, x, y); store(o, z); }
void h_fe_invert(uint8_t *o, const uint8_t *a) { fe x, z; load(x, a); fe_invert(z, x); store(o, z); }
void h_fe_pow_p58(uint8_t *o, const uint8_t *a) { fe x, z; load(x, a); fe_pow_p58(z, x); store(o, z); }
// chained stress: ((a*b - a)^2 + b) * (a - b) ... exercises the scale bookkeeping
void h_fe_chain(uint8_t *o, const uint8_t *a, const uint8_t *b)
{
    fe x, y, t, u, v; load(x, a); load(y, b);
    fe_mul(t, x, y); fe_sub(u, t, x); fe_carry(u, u); fe_sq(u, u); fe_add(u, u, y);
    fe_sub(v, x, y); fe_mul(t, v, u); store(o, t);
}
int h_decompress(uint8_t *ox, uint8_t *oy, const uint8_t *s)
{
    uint32_t w[8]; memcpy(w, s, 32); fe x, y; int ok = (int)ge_decompress_affine(x, y, w);
    store(ox, x); store(oy, y); return ok;
}
// out = compress( k1 * P + ... ) helpers: expose point ops on compressed inputs
static int load_point(ge_p3 &p, const uint8_t *s)
{
    uint32_t w[8]; memcpy(w, s, 32); fe x, y; if (!ge_decompress_affine(x, y, w)) return 0;
    p.X = x; p.Y = y; fe_1(p.Z); fe_mul(p.T, x, y); return 1;
}
static void store_point(uint8_t *s, const ge_p3 &p) { uint32_t w[8]; ge_compress(w, p); memcpy(s, w, 32); }
// r = 2^k * (P + Q) - Q + Q(affine niels) ... returns several results for comparison
int h_point_ops(uint8_t *o_add, uint8_t *o_sub, uint8_t *o_madd, uint8_t *o_msub, uint8_t *o_dbl,
                uint8_t *o_pow2k, const uint8_t *P, const uint8_t *Q, int k)
{
    ge_p3 p, q, r;
    if (!load_point(p, P) || !load_point(q, Q)) return 0;
    // scale p's Z to make it properly projective: p = 2*(p) - p ... use add chain instead
    ge_add(r, p, q); ge_add(r, r, q); ge_pniels nq; ge_p3_to_pniels(nq, q); ge_padd(r, r, nq, 1);   // p + q
    store_point(o_add, r);
    ge_padd(r, p, nq, 1); store_point(o_sub, r);                                                  // p - q
    ge_niels n; ge_affine_to_niels(n, q.X, q.Y);
    ge_p3 pp; ge_dbl(pp, p); ge_padd(pp, pp, nq, 0);            // pp = 2p + q (Z != 1)
    ge_madd(r, pp, n, 0); store_point(o_madd, r);               // 2p + 2q
    ge_madd(r, pp, n, 1); store_point(o_msub, r);               // 2p
    ge_dbl(r, pp); store_point(o_dbl, r);                       // 4p + 2q
    ge_mul_by_pow_2(r, pp, k); store_point(o_pow2k, r);         // 2^k (2p + q)
    // packed round trips
    ge_niels_packed np; ge_niels_pack(np, n); ge_niels n2; ge_niels_unpack(n2, np);
    ge_p3 r2; ge_madd(r2, pp, n2, 0); uint8_t chk[32]; store_point(chk, r2);
    if (memcmp(chk, o_madd, 32)) return -1;
    ge_pniels_packed pnp; ge_pniels_pack(pnp, nq); ge_pniels nq2; ge_pniels_unpack(nq2, pnp);
    ge_padd(r2, p, nq2, 1); store_point(chk, r2);
    if (memcmp(chk, o_sub, 32)) return -2;
    return 1;
}
void h_limbs51_roundtrip(uint64_t *out, const uint64_t *in)
{
    for (int c = 0; c < 4; c++) { fe f; fe_from_limbs51(f, in + 5 * c); fe_to_limbs51(out + 5 * c, f); }
}
int h_is_identity_of_diff(const uint8_t *P)
{
    ge_p3 p, r; if (!load_point(p, P)) return -1;
    ge_pniels n; ge_p3_to_pniels(n, p); ge_dbl(r, p); ge_padd(r, r, n, 1); ge_padd(r, r, n, 1);
    return (int)ge_is_identity(r);
}
}
extern "C" {
int h_ristretto_roundtrip(uint8_t *out, uint8_t *out_dbl, const uint8_t *in)
{
    uint32_t w[8]; memcpy(w, in, 32);
    ge_p3 p;
    if (!ristretto_decompress(p, w)) return 0;
    uint32_t o[8]; ristretto_compress(o, p); memcpy(out, o, 32);
    ge_p3 q; ge_dbl(q, p); ge_add(q, q, p);           // 3P with Z != 1
    ristretto_compress(o, q); memcpy(out_dbl, o, 32);
    return 1;
}
}
#include "../../curve25519_dalek_b200/csrc/ge64.cuh"
extern "C" {
// FP64-field model (host emulation of the exact arithmetic): product, and a chain of mixed additions
void h_fe64_mul(uint8_t *o, const uint8_t *a, const uint8_t *b)
{
    fe x, y, z; load(x, a); load(y, b);
    fe64 X, Y, Z; fe64_from_fe(X, x); fe64_from_fe(Y, y);
    fe64_mul(Z, X, Y); fe64_mul(Z, Z, Y);
    fe64 S; fe64_add(S, Z, X); fe64_sub(S, S, Y); fe64_mul(Z, S, X);      // (a b^2 + a - b) a
    fe64_to_fe(z, Z); store(o, z);
}
// x^((p-5)/8) with the addition chain on the FP64 field (fe64_sq / fe64_mul host models, bound asserts on)
void h_fe64_pow_p58(uint8_t *o, const uint8_t *a)
{
    fe x, z; load(x, a);
    fe_pow_p58_f64(z, x); store(o, z);
}
void h_fe64_invert(uint8_t *o, const uint8_t *a)
{
    fe x, z; load(x, a);
    fe_invert_f64(z, x); store(o, z);
}
void h_fe64_sq(uint8_t *o, const uint8_t *a, const uint8_t *b)
{
    fe x, y, z; load(x, a); load(y, b);
    fe64 X, Y, S; fe64_from_fe(X, x); fe64_from_fe(Y, y);
    fe64_sq(S, X); fe64_sq(S, S);                  // a^4
    fe64_mul(S, S, Y); fe64_sq(S, S);              // (a^4 b)^2
    fe64_to_fe(z, S); store(o, z);
}
// acc = sum_k (+/-) Q_k over `count` compressed points, alternating through madd / padd; returns compress(acc)
int h_ge64_chain(uint8_t *out, const uint8_t *pts, const uint8_t *negs, int count)
{
    ge64_p3 acc; ge64_identity(acc);
    for (int k = 0; k < count; k++) {
        ge_p3 q; if (!load_point(q, pts + 32 * k)) return 0;
        if (k & 1) {
            ge_niels n; ge_affine_to_niels(n, q.X, q.Y);
            ge_niels_packed pk; ge_niels_pack(pk, n);
            ge64_niels n64; ge64_niels_unpack(n64, pk);
            ge64_madd(acc, acc, n64, negs[k]);
        } else {
            ge_p3 q3; ge_dbl(q3, q); ge_add(q3, q3, q);                    // 3q, Z != 1
            ge_pniels n; ge_p3_to_pniels(n, q3);
            ge_pniels_packed pk; ge_pniels_pack(pk, n);
            ge64_pniels n64; ge64_pniels_unpack(n64, pk);
            ge64_padd(acc, acc, n64, negs[k]);
        }
    }
    ge_p3 r; ge64_to_p3(r, acc);
    store_point(out, r);
    return 1;
}
// Q = 3P; 2^k * Q with ge64_dbl, then + Q through ge64_padd: returns compress((2^k + 1) * 3P)
int h_ge64_dbl_chain(uint8_t *out, const uint8_t *pt, int k)
{
    ge_p3 q; if (!load_point(q, pt)) return 0;
    ge_p3 q3; ge_dbl(q3, q); ge_add(q3, q3, q);                          // 3q, Z != 1
    ge64_p3 acc; ge64_from_p3(acc, q3);
    for (int i = 0; i < k; i++) ge64_dbl(acc, acc);
    ge_pniels n; ge_p3_to_pniels(n, q3);
    ge_pniels_packed pk; ge_pniels_pack(pk, n);
    ge64_pniels n64; ge64_pniels_unpack(n64, pk);
    ge64_padd(acc, acc, n64, 0);
    ge_p3 r; ge64_to_p3(r, acc);
    store_point(out, r);
    return 1;
}
}
