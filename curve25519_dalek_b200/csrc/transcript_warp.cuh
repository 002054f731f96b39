// transcript_warp.cuh -- ONE Merlin transcript of ed25519_dalek::verify_batch computed by ONE WARP.
//
// The reference draws the batch coefficients z_i from a single Merlin transcript (ed25519-dalek/src/batch.rs:168-222,
// src/batch/transcript.rs:54-207; STROBE-128 over Keccak-f[1600]): append every hram, append every s, fork an RNG with a
// zero witness, then squeeze 16 bytes per signature.  It is a strictly sequential sponge -- 1.73 permutations per
// signature -- so what can be parallelised is the inside of each step:
//
//   * the 25 lanes of the Keccak state live in 25 lanes of the warp; theta, pi and chi exchange lanes with shuffles
//     (18 shuffles and ~45 instructions per round instead of ~200 dependent instructions on one thread);
//   * the bytes absorbed into a 166-byte rate block are a closed-form function of their position in the transcript's
//     byte stream (frames of 76 bytes per hram, 45 bytes per s, STROBE's `old_begin` bookkeeping bytes included), so
//     21 lanes assemble their 8 bytes of the block independently.
//
// The state after `Transcript::new(b"ed25519 batch verification")` (transcript.rs:54-61) does not depend on the batch
// and is a constant (MERLIN_PREFIX_*; tests/test_w4_host.py recomputes it with the oracle).
//
// Written against the lane primitives of warp4_f64.cuh so that tests/host can run it on an emulated warp.
#pragma once
#include <stdint.h>

#include "warp4_f64.cuh"

#define TW_RATE 166u
#define MERLIN_PREFIX_POS 54u            // in-block position after the dom-sep message
#define MERLIN_PREFIX_POS_BEGIN 27u      // STROBE pos_begin at that point

FE_HD uint64_t merlin_prefix_lane(uint32_t lane)
{
    // Keccak state after strobe128::new("Merlin v1.0") and append_message("dom-sep", "ed25519 batch verification")
    switch (lane) {
    case 0: return 0xb43c918aea5b7f9cULL; case 1: return 0x072764650d0dd10aULL; case 2: return 0xf63c65302f6a61b3ULL;
    case 3: return 0x4987089420e43b73ULL; case 4: return 0x719d6f040e512ee6ULL; case 5: return 0xfd75877347016aeaULL;
    case 6: return 0x123343d9953e0e41ULL; case 7: return 0xf6ac24a6e892cc93ULL; case 8: return 0xfbbb22e39500b6e1ULL;
    case 9: return 0x7dfe9569b2e545c8ULL; case 10: return 0x9858ffd17413847cULL; case 11: return 0x7372066b63e02ec9ULL;
    case 12: return 0x53030739602ac921ULL; case 13: return 0x05b0b7921bbbcc49ULL; case 14: return 0x887ebcce7fa88f7eULL;
    case 15: return 0x34bc04ae45cb6f65ULL; case 16: return 0x5017d979beaebecaULL; case 17: return 0x4d5066b913bfe8c0ULL;
    case 18: return 0x6588dd6572594313ULL; case 19: return 0xd5209bcc0914f9adULL; case 20: return 0x99b6971f044474f4ULL;
    case 21: return 0xd07ba81ee9defbddULL; case 22: return 0xe9965aa72db0f89bULL; case 23: return 0x6e4ebb655b7ff047ULL;
    case 24: return 0xf6fbd9bf6aa1fafeULL; default: return 0;
    }
}

FE_HD uint64_t tw_keccak_rc(int round)
{
    switch (round) {
    case 0: return 0x0000000000000001ULL; case 1: return 0x0000000000008082ULL; case 2: return 0x800000000000808aULL;
    case 3: return 0x8000000080008000ULL; case 4: return 0x000000000000808bULL; case 5: return 0x0000000080000001ULL;
    case 6: return 0x8000000080008081ULL; case 7: return 0x8000000000008009ULL; case 8: return 0x000000000000008aULL;
    case 9: return 0x0000000000000088ULL; case 10: return 0x0000000080008009ULL; case 11: return 0x000000008000000aULL;
    case 12: return 0x000000008000808bULL; case 13: return 0x800000000000008bULL; case 14: return 0x8000000000008089ULL;
    case 15: return 0x8000000000008003ULL; case 16: return 0x8000000000008002ULL; case 17: return 0x8000000000000080ULL;
    case 18: return 0x000000000000800aULL; case 19: return 0x800000008000000aULL; case 20: return 0x8000000080008081ULL;
    case 21: return 0x8000000000008080ULL; case 22: return 0x0000000080000001ULL; default: return 0x8000000080008008ULL;
    }
}

// rho offsets r[x + 5 y]
FE_HD uint32_t tw_rho(uint32_t lane)
{
    const uint8_t r[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    return r[lane % 25];
}

W4_DEV uint64_t tw_shfl64(uint64_t v, int src)
{
    const uint32_t lo = w4_shfl((uint32_t)v, src), hi = w4_shfl((uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// where each lane finds its operands (x = lane % 5, y = lane / 5; lanes 25..31 mirror lanes 0..6 and are ignored)
struct tw_lanes { int c1, c2, c3, c4, dm, dp, pi, x1, x2; uint32_t rot; };
W4_DEV void tw_lanes_init(tw_lanes &t, uint32_t lane)
{
    const int l = (int)(lane % 25), x = l % 5, y = l / 5;
    t.c1 = x + 5 * ((y + 1) % 5); t.c2 = x + 5 * ((y + 2) % 5); t.c3 = x + 5 * ((y + 3) % 5); t.c4 = x + 5 * ((y + 4) % 5);
    t.dm = (x + 4) % 5 + 5 * y; t.dp = (x + 1) % 5 + 5 * y;
    t.pi = (x + 3 * y) % 5 + 5 * x;                       // B[x', y'] = rot(A[x, y]) with x' = y, y' = 2x + 3y
    t.x1 = (x + 1) % 5 + 5 * y; t.x2 = (x + 2) % 5 + 5 * y;
    t.rot = tw_rho((uint32_t)l);
}

// Keccak-f[1600] on a state spread over lanes 0..24 (one 64-bit lane each)
W4_DEV uint64_t tw_keccak_f(uint64_t a, const tw_lanes &t, uint32_t lane)
{
#if FE64_DEV
#pragma unroll 1
#endif
    for (int round = 0; round < 24; round++) {
        // theta
        const uint64_t c = a ^ tw_shfl64(a, t.c1) ^ tw_shfl64(a, t.c2) ^ tw_shfl64(a, t.c3) ^ tw_shfl64(a, t.c4);
        const uint64_t cm = tw_shfl64(c, t.dm), cp = tw_shfl64(c, t.dp);
        a ^= cm ^ ((cp << 1) | (cp >> 63));
        // rho (own offset), then pi (gather)
        const uint64_t r = t.rot ? ((a << t.rot) | (a >> (64 - t.rot))) : a;
        const uint64_t b = tw_shfl64(r, t.pi);
        // chi
        const uint64_t b1 = tw_shfl64(b, t.x1), b2 = tw_shfl64(b, t.x2);
        a = b ^ (~b1 & b2);
        if (lane == 0) a ^= tw_keccak_rc(round);
    }
    return a;
}

// ---- the transcript's absorbed byte stream, position u = 0 right after the prefix --------------------------------
//   u in [0, 76 n)          hram frames:  ob 0x12 'h' 'r' 'a' 'm' 64 0 0 0 ob 0x02 <64 bytes>
//   u in [76 n, 121 n)      s frames:     ob 0x12 's' 'i' 'g' '.' 's' 32 0 0 0 ob 0x02 <32 bytes>
//   u in [121 n, 121 n + 7) rng fork:     ob 0x12 'r' 'n' 'g' ob 0x06                     (transcript.rs:157-173)
// `ob` = STROBE's old_begin: in-block position + 1 of the previous operation's start if it lies in the same rate
// block, else 0.
struct tw_stream { const uint32_t *hrams; const uint32_t *sigs; uint64_t n; };

FE_HD uint64_t tw_last_op_before(const tw_stream &s, uint64_t x)        // largest operation start < x   (x >= 1)
{
    const uint64_t nb = 76 * s.n, nc = 121 * s.n;
    if (x <= nb) { const uint64_t f = (x - 1) / 76, r = (x - 1) % 76; return f * 76 + (r >= 10 ? 10 : 0); }
    if (x <= nc) { const uint64_t y = x - nb, f = (y - 1) / 45, r = (y - 1) % 45; return nb + f * 45 + (r >= 11 ? 11 : 0); }
    return nc + ((x - nc - 1) >= 5 ? 5 : 0);
}
FE_HD uint32_t tw_old_begin(const tw_stream &s, uint64_t u)             // the `ob` byte of the operation starting at u
{
    if (u == 0) return MERLIN_PREFIX_POS_BEGIN;                         // same block as the prefix's last operation
    const uint64_t prev = tw_last_op_before(s, u);
    const uint64_t bu = (u + MERLIN_PREFIX_POS) / TW_RATE, bp = (prev + MERLIN_PREFIX_POS) / TW_RATE;
    return bu == bp ? (uint32_t)((prev + MERLIN_PREFIX_POS) % TW_RATE) + 1u : 0u;
}
FE_HD uint32_t tw_byte_at(const tw_stream &s, uint64_t u)
{
    const uint64_t nb = 76 * s.n, nc = 121 * s.n;
    if (u < nb) {
        const uint64_t f = u / 76; const uint32_t r = (uint32_t)(u % 76);
        if (r >= 12) { const uint32_t k = r - 12; return (s.hrams[16 * f + (k >> 2)] >> (8 * (k & 3))) & 0xffu; }
        switch (r) {
        case 0: case 10: return tw_old_begin(s, u);
        case 1: return 0x12; case 2: return 'h'; case 3: return 'r'; case 4: return 'a'; case 5: return 'm';
        case 6: return 64; case 11: return 0x02; default: return 0;
        }
    }
    if (u < nc) {
        const uint64_t y = u - nb, f = y / 45; const uint32_t r = (uint32_t)(y % 45);
        if (r >= 13) { const uint32_t k = r - 13; return (s.sigs[16 * f + 8 + (k >> 2)] >> (8 * (k & 3))) & 0xffu; }
        switch (r) {
        case 0: case 11: return tw_old_begin(s, u);
        case 1: return 0x12; case 2: return 's'; case 3: return 'i'; case 4: return 'g'; case 5: return '.'; case 6: return 's';
        case 7: return 32; case 12: return 0x02; default: return 0;
        }
    }
    switch ((uint32_t)(u - nc)) {
    case 0: case 5: return tw_old_begin(s, u);
    case 1: return 0x12; case 2: return 'r'; case 3: return 'n'; case 4: return 'g'; default: return 0x06;
    }
}

// The coefficients z_0 .. z_{n-1} (four little-endian words each) of ONE transcript over signatures 0 .. n-1:
// hrams = n x 16 words (SHA-512(R || A || M)), sigs = n x 16 words (s in words 8..15).  All 32 lanes call this.
W4_DEV void merlin_zs_warp(const uint32_t *hrams, const uint32_t *sigs, uint64_t n, uint32_t *zs)
{
    const uint32_t lane = w4_lane();
    tw_lanes tl; tw_lanes_init(tl, lane);
    const tw_stream s = {hrams, sigs, n};
    uint64_t a = lane < 25 ? merlin_prefix_lane(lane) : 0;
    const uint64_t T = 121 * n + 7;                                     // bytes absorbed after the prefix
    // ---- absorb: rate blocks k = 0, 1, ...; block k holds stream positions [k R - pos0, (k + 1) R - pos0)
    for (uint64_t k = 0;; k++) {
        const uint64_t base = k * TW_RATE;                              // in-block position p <-> u = base + p - pos0
        const uint64_t end = (base + TW_RATE - MERLIN_PREFIX_POS) < T ? (base + TW_RATE - MERLIN_PREFIX_POS) : T;
        const bool full = end == base + TW_RATE - MERLIN_PREFIX_POS;    // the block fills up: run_f at pos = R
        if (lane < 21) {
            uint64_t w = 0;
#if FE64_DEV
#pragma unroll 1
#endif
            for (uint32_t b = 0; b < 8; b++) {
                const uint32_t p = 8 * lane + b;
                if (p >= TW_RATE || base + p < MERLIN_PREFIX_POS) continue;
                const uint64_t u = base + p - MERLIN_PREFIX_POS;
                if (u < end) w |= (uint64_t)tw_byte_at(s, u) << (8 * b);
            }
            a ^= w;
        }
        // run_f (strobe: st[pos] ^= pos_begin, st[pos + 1] ^= 0x04, st[R + 1] ^= 0x80), pos = R for a full block
        const uint32_t pos = full ? TW_RATE : (uint32_t)((end + MERLIN_PREFIX_POS) % TW_RATE);
        if (full || pos != 0) {
            const uint64_t last = tw_last_op_before(s, end);            // operations are at most 66 bytes apart: it lies in this block
            const uint32_t pb = (uint32_t)((last + MERLIN_PREFIX_POS) % TW_RATE) + 1u;
            if (lane == (pos >> 3)) a ^= (uint64_t)pb << (8 * (pos & 7));
            if (lane == ((pos + 1) >> 3)) a ^= (uint64_t)0x04 << (8 * ((pos + 1) & 7));
            if (lane == 20) a ^= (uint64_t)0x80 << 56;                  // byte R + 1 = 167
            a = tw_keccak_f(a, tl, lane);
        }
        if (end == T) break;
    }
    // ---- KEY with 32 zero bytes (ZeroRng witness, batch.rs:49-76): state bytes 0..31 overwritten, pos = 32
    if (lane < 4) a = 0;
    // ---- per signature: meta_ad(16u32 LE), prf(16)   (transcript.rs:200-206)
    uint32_t P = 32;
    for (uint64_t i = 0; i < n; i++) {
        // bytes P .. P+7: 0, 0x12, 16, 0, 0, 0, P+1, 0x07; run_f at P+8: st[P+8] ^= P+7, st[P+9] ^= 0x04, st[167] ^= 0x80
        const uint64_t w0 = ((uint64_t)0x12 << 8) | ((uint64_t)16 << 16) | ((uint64_t)(P + 1) << 48) | ((uint64_t)0x07 << 56);
        const uint64_t w1 = (uint64_t)(P + 7) | ((uint64_t)0x04 << 8);
        if (lane == (P >> 3)) a ^= w0;
        if (lane == (P >> 3) + 1) a ^= w1;
        if (lane == 20) a ^= (uint64_t)0x80 << 56;
        a = tw_keccak_f(a, tl, lane);
        if (lane < 2) { zs[4 * i + 2 * lane] = (uint32_t)a; zs[4 * i + 2 * lane + 1] = (uint32_t)(a >> 32); a = 0; }
        P = 16;
    }
}
