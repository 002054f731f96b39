/*
 * hash.c -- SHA-512, Keccak-f[1600], STROBE-128 subset, Merlin transcript.
 * TEST INFRASTRUCTURE (oracle).
 *
 * Third-party algorithms not present under /root/reference (Cargo.lock pins):
 *   sha2 0.11.0      -> SHA-512, FIPS 180-4            (call sites E/batch.rs:185-189)
 *   keccak 0.2.0     -> Keccak-f[1600], FIPS 202
 *   strobe-rs 0.13.0 -> STROBE v1.0.2, 128-bit security (E/batch/transcript.rs:56,71-73,167-168,202-203)
 * They are restated from the published specifications.  SHA-512 and Keccak-f are
 * pinned against hashlib in tests; the STROBE framing is pinned against the public
 * merlin "Conformance Test Protocol" vector; the reference tree itself holds no
 * golden transcript output ("parity unpinned" for z_i values, see oracle.h).
 */
#include "oracle.h"
#include <string.h>

/* ---------------- SHA-512 ---------------- */
static const uint64_t K512[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL,
    0x3956c25bf348b538ULL, 0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL,
    0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, 0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL,
    0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, 0xc19bf174cf692694ULL,
    0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL,
    0x983e5152ee66dfabULL, 0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL,
    0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, 0x06ca6351e003826fULL, 0x142929670a0e6e70ULL,
    0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, 0x53380d139d95b3dfULL,
    0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL,
    0xd192e819d6ef5218ULL, 0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL,
    0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, 0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL,
    0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, 0x682e6ff3d6b2b8a3ULL,
    0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL,
    0xca273eceea26619cULL, 0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL,
    0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, 0x113f9804bef90daeULL, 0x1b710b35131c471bULL,
    0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, 0x431d67c49c100d4cULL,
    0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};

static inline uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

static void sha512_block(uint64_t h[8], const uint8_t *p)
{
    uint64_t w[80];
    for (int i = 0; i < 16; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | p[8 * i + j];
        w[i] = v;
    }
    for (int i = 16; i < 80; i++) {
        uint64_t s0 = ror64(w[i - 15], 1) ^ ror64(w[i - 15], 8) ^ (w[i - 15] >> 7);
        uint64_t s1 = ror64(w[i - 2], 19) ^ ror64(w[i - 2], 61) ^ (w[i - 2] >> 6);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 80; i++) {
        uint64_t S1 = ror64(e, 14) ^ ror64(e, 18) ^ ror64(e, 41);
        uint64_t ch = (e & f) ^ (~e & g);
        uint64_t t1 = hh + S1 + ch + K512[i] + w[i];
        uint64_t S0 = ror64(a, 28) ^ ror64(a, 34) ^ ror64(a, 39);
        uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint64_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

void sha512_init(sha512_ctx *c)
{
    static const uint64_t iv[8] = {
        0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
        0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    memcpy(c->h, iv, sizeof iv);
    c->buflen = 0; c->total = 0;
}

void sha512_update(sha512_ctx *c, const uint8_t *msg, size_t len)
{
    c->total += len;
    while (len) {
        size_t take = 128 - c->buflen; if (take > len) take = len;
        memcpy(c->buf + c->buflen, msg, take);
        c->buflen += take; msg += take; len -= take;
        if (c->buflen == 128) { sha512_block(c->h, c->buf); c->buflen = 0; }
    }
}

void sha512_final(sha512_ctx *c, uint8_t out[64])
{
    uint64_t bits = c->total * 8;
    uint8_t pad = 0x80;
    sha512_update(c, &pad, 1);
    uint8_t z = 0;
    while (c->buflen != 112) sha512_update(c, &z, 1);
    uint8_t lenb[16] = {0};
    for (int i = 0; i < 8; i++) lenb[15 - i] = (uint8_t)(bits >> (8 * i));
    sha512_update(c, lenb, 16);
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(c->h[i] >> (56 - 8 * j));
}

void sha512(uint8_t out[64], const uint8_t *msg, size_t len)
{
    sha512_ctx c; sha512_init(&c); sha512_update(&c, msg, len); sha512_final(&c, out);
}

/* ---------------- Keccak-f[1600] (FIPS 202) ---------------- */
static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

void keccak_f1600(uint64_t st[25])
{
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39,
                                41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; round++) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rol64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) st[i] ^= D[i % 5];
        /* rho + pi: B[y, 2x+3y] = rot(A[x,y]) ; index = x + 5y */
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++)
                B[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(st[x + 5 * y], ROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++)
                st[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        st[0] ^= RC[round];
    }
}

/* ---------------- STROBE-128 (subset: AD, meta-AD, PRF, KEY) ---------------- */
#define STROBE_R 166
#define FLAG_I 1
#define FLAG_A 2
#define FLAG_C 4
#define FLAG_T 8
#define FLAG_M 16
#define FLAG_K 32

static void strobe_run_f(strobe128 *s)
{
    s->st[s->pos] ^= s->pos_begin;
    s->st[s->pos + 1] ^= 0x04;
    s->st[STROBE_R + 1] ^= 0x80;
    uint64_t lanes[25];
    for (int i = 0; i < 25; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v |= (uint64_t)s->st[8 * i + j] << (8 * j);
        lanes[i] = v;
    }
    keccak_f1600(lanes);
    for (int i = 0; i < 25; i++)
        for (int j = 0; j < 8; j++) s->st[8 * i + j] = (uint8_t)(lanes[i] >> (8 * j));
    s->pos = 0; s->pos_begin = 0;
}

static void strobe_absorb(strobe128 *s, const uint8_t *d, size_t len)
{
    for (size_t i = 0; i < len; i++) {
        s->st[s->pos] ^= d[i];
        if (++s->pos == STROBE_R) strobe_run_f(s);
    }
}

static void strobe_overwrite(strobe128 *s, const uint8_t *d, size_t len)
{
    for (size_t i = 0; i < len; i++) {
        s->st[s->pos] = d[i];
        if (++s->pos == STROBE_R) strobe_run_f(s);
    }
}

static void strobe_squeeze(strobe128 *s, uint8_t *d, size_t len)
{
    for (size_t i = 0; i < len; i++) {
        d[i] = s->st[s->pos];
        s->st[s->pos] = 0;
        if (++s->pos == STROBE_R) strobe_run_f(s);
    }
}

static void strobe_begin_op(strobe128 *s, uint8_t flags, int more)
{
    if (more) return; /* continuing the current operation (flags must match) */
    uint8_t old_begin = s->pos_begin;
    s->pos_begin = (uint8_t)(s->pos + 1);
    s->cur_flags = flags;
    uint8_t hdr[2] = {old_begin, flags};
    strobe_absorb(s, hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && s->pos != 0) strobe_run_f(s);
}

void strobe128_new(strobe128 *s, const uint8_t *proto, size_t len)
{
    memset(s, 0, sizeof *s);
    static const uint8_t init[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(s->st, init, 6);
    memcpy(s->st + 6, "STROBEv1.0.2", 12);
    uint64_t lanes[25];
    for (int i = 0; i < 25; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v |= (uint64_t)s->st[8 * i + j] << (8 * j);
        lanes[i] = v;
    }
    keccak_f1600(lanes);
    for (int i = 0; i < 25; i++)
        for (int j = 0; j < 8; j++) s->st[8 * i + j] = (uint8_t)(lanes[i] >> (8 * j));
    s->pos = 0; s->pos_begin = 0; s->cur_flags = 0;
    strobe128_meta_ad(s, proto, len, 0);
}

void strobe128_meta_ad(strobe128 *s, const uint8_t *d, size_t len, int more)
{ strobe_begin_op(s, FLAG_M | FLAG_A, more); strobe_absorb(s, d, len); }
void strobe128_ad(strobe128 *s, const uint8_t *d, size_t len, int more)
{ strobe_begin_op(s, FLAG_A, more); strobe_absorb(s, d, len); }
void strobe128_prf(strobe128 *s, uint8_t *d, size_t len, int more)
{ strobe_begin_op(s, FLAG_I | FLAG_A | FLAG_C, more); strobe_squeeze(s, d, len); }
void strobe128_key(strobe128 *s, const uint8_t *d, size_t len, int more)
{ strobe_begin_op(s, FLAG_A | FLAG_C, more); strobe_overwrite(s, d, len); }

/* ---------------- Merlin transcript, E/batch/transcript.rs ---------------- */
static void le32(uint8_t o[4], size_t x) { for (int i = 0; i < 4; i++) o[i] = (uint8_t)(x >> (8 * i)); }

/* transcript.rs:54-61 */
void merlin_new(merlin_transcript *t, const uint8_t *label, size_t len)
{
    strobe128_new(&t->s, (const uint8_t *)"Merlin v1.0", 11);     /* E/batch.rs:44 */
    merlin_append_message(t, (const uint8_t *)"dom-sep", 7, label, len);
}

/* transcript.rs:69-74 */
void merlin_append_message(merlin_transcript *t, const uint8_t *label, size_t llen,
                           const uint8_t *msg, size_t mlen)
{
    uint8_t dl[4]; le32(dl, mlen);
    strobe128_meta_ad(&t->s, label, llen, 0);
    strobe128_meta_ad(&t->s, dl, 4, 1);
    strobe128_ad(&t->s, msg, mlen, 0);
}

/* transcript.rs:83-88 */
void merlin_challenge_bytes(merlin_transcript *t, const uint8_t *label, size_t llen,
                            uint8_t *dest, size_t dlen)
{
    uint8_t dl[4]; le32(dl, dlen);
    strobe128_meta_ad(&t->s, label, llen, 0);
    strobe128_meta_ad(&t->s, dl, 4, 1);
    strobe128_prf(&t->s, dest, dlen, 0);
}

/* transcript.rs:96-100 build_rng + :157-173 finalize with ZeroRng (E/batch.rs:49-76):
 * random_bytes stays [0u8; 32] */
void merlin_rng_finalize_zero(merlin_transcript *rng, const merlin_transcript *t)
{
    uint8_t zero[32] = {0};
    *rng = *t;
    strobe128_meta_ad(&rng->s, (const uint8_t *)"rng", 3, 0);
    strobe128_key(&rng->s, zero, 32, 0);
}

/* transcript.rs:200-206 */
void merlin_rng_fill(merlin_transcript *rng, uint8_t *dest, size_t dlen)
{
    uint8_t dl[4]; le32(dl, dlen);
    strobe128_meta_ad(&rng->s, dl, 4, 0);
    strobe128_prf(&rng->s, dest, dlen, 0);
}
