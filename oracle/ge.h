// ORACLE — test infrastructure only (see fe.h header).
// Ristretto255 over the twisted Edwards curve -x^2 + y^2 = 1 + d x^2 y^2 in extended coordinates.
// Restates curve25519-dalek 1.2.3 EdwardsPoint/RistrettoPoint/CompressedRistretto (SURVEY.md §2.2 U3,
// Appendix A: decompress / compress / from_uniform_bytes / unified add). Pinned against RFC 9496
// vectors and libsodium's crypto_core_ristretto255_* in tests/test_oracle_primitives.py.
#pragma once
#include "fe.h"
#include "sc.h"

namespace orc {

struct ge {
    fe X, Y, Z, T;
};

static inline ge ge_identity() { return ge{fe_zero(), fe_one(), fe_one(), fe_zero()}; }

// unified addition (Appendix A "Unified add"), complete on the prime-order subgroup and its cosets
static inline ge ge_add(const ge &p, const ge &q) {
    const fe_consts &K = fe_constants();
    fe A = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
    fe B = fe_mul(fe_add(p.Y, p.X), fe_add(q.Y, q.X));
    fe C = fe_mul(fe_mul(p.T, q.T), K.d2);
    fe D = fe_mul(p.Z, q.Z);
    D = fe_add(D, D);
    fe E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
    return ge{fe_mul(E, F), fe_mul(G, H), fe_mul(F, G), fe_mul(E, H)};
}
static inline ge ge_neg(const ge &p) { return ge{fe_neg(p.X), p.Y, p.Z, fe_neg(p.T)}; }
static inline ge ge_sub(const ge &p, const ge &q) { return ge_add(p, ge_neg(q)); }

// dedicated doubling (dbl-2008-hwcd, a = -1)
static inline ge ge_dbl(const ge &p) {
    fe A = fe_sq(p.X), B = fe_sq(p.Y);
    fe C = fe_sq(p.Z);
    C = fe_add(C, C);
    fe D = fe_neg(A);                                   // a*A
    fe xy = fe_add(p.X, p.Y);
    fe E = fe_sub(fe_sub(fe_sq(xy), A), B);
    fe G = fe_add(D, B), F = fe_sub(G, C), H = fe_sub(D, B);
    return ge{fe_mul(E, F), fe_mul(G, H), fe_mul(F, G), fe_mul(E, H)};
}

// Ristretto equality: X1*Y2 == Y1*X2 or Y1*Y2 == X1*X2
static inline bool ge_eq(const ge &p, const ge &q) {
    return fe_eq(fe_mul(p.X, q.Y), fe_mul(p.Y, q.X)) || fe_eq(fe_mul(p.Y, q.Y), fe_mul(p.X, q.X));
}

static inline void ge_compress(uint8_t out[32], const ge &p) {
    const fe_consts &K = fe_constants();
    fe X = p.X, Y = p.Y;
    fe u1 = fe_mul(fe_add(p.Z, Y), fe_sub(p.Z, Y));
    fe u2 = fe_mul(X, Y);
    fe I;
    fe_sqrt_ratio_i(I, fe_one(), fe_mul(u1, fe_sq(u2)), K.sqrt_m1);
    fe d1 = fe_mul(I, u1), d2 = fe_mul(I, u2);
    fe z_inv = fe_mul(fe_mul(d1, d2), p.T);
    fe d_inv;
    if (fe_isneg(fe_mul(p.T, z_inv))) {
        fe nx = fe_mul(Y, K.sqrt_m1), ny = fe_mul(X, K.sqrt_m1);
        X = nx; Y = ny;
        d_inv = fe_mul(d1, K.invsqrt_a_minus_d);
    } else {
        d_inv = d2;
    }
    if (fe_isneg(fe_mul(X, z_inv))) Y = fe_neg(Y);
    fe s = fe_abs(fe_mul(d_inv, fe_sub(p.Z, Y)));
    fe_tobytes(out, s);
}

static inline bool ge_decompress(ge &out, const uint8_t in[32]) {
    const fe_consts &K = fe_constants();
    fe s = fe_frombytes(in);
    uint8_t chk[32];
    fe_tobytes(chk, s);
    if (memcmp(chk, in, 32) != 0) return false;   // non-canonical (>= p or bit 255 set)
    if (in[0] & 1) return false;                  // negative
    fe one = fe_one();
    fe ss = fe_sq(s);
    fe u1 = fe_sub(one, ss), u2 = fe_add(one, ss);
    fe u2s = fe_sq(u2);
    fe v = fe_sub(fe_neg(fe_mul(K.d, fe_sq(u1))), u2s);
    fe I;
    bool ok = fe_sqrt_ratio_i(I, one, fe_mul(v, u2s), K.sqrt_m1);
    fe Dx = fe_mul(I, u2);
    fe Dy = fe_mul(fe_mul(I, Dx), v);
    fe x = fe_abs(fe_mul(fe_add(s, s), Dx));
    fe y = fe_mul(u1, Dy);
    fe t = fe_mul(x, y);
    if (!ok || fe_isneg(t) || fe_iszero(y)) return false;
    out = ge{x, y, one, t};
    return true;
}

// Elligator map of one field element (Appendix A MAP)
static inline ge ge_elligator(const fe &r0) {
    const fe_consts &K = fe_constants();
    fe one = fe_one();
    fe r = fe_mul(K.sqrt_m1, fe_sq(r0));
    fe u = fe_mul(fe_add(r, one), K.one_minus_d_sq);
    fe minus_one = fe_neg(one);
    fe v = fe_mul(fe_sub(minus_one, fe_mul(r, K.d)), fe_add(r, K.d));
    fe s;
    bool sq = fe_sqrt_ratio_i(s, u, v, K.sqrt_m1);
    fe sp = fe_neg(fe_abs(fe_mul(s, r0)));
    fe c;
    if (!sq) { s = sp; c = r; } else { c = minus_one; }
    fe N = fe_sub(fe_mul(fe_mul(c, fe_sub(r, one)), K.d_minus_one_sq), v);
    fe ss = fe_sq(s);
    fe w0 = fe_mul(fe_add(s, s), v);
    fe w1 = fe_mul(N, K.sqrt_ad_minus_one);
    fe w2 = fe_sub(one, ss), w3 = fe_add(one, ss);
    return ge{fe_mul(w0, w3), fe_mul(w2, w1), fe_mul(w1, w3), fe_mul(w0, w2)};
}

// RistrettoPoint::from_uniform_bytes (= libsodium crypto_core_ristretto255_from_hash)
static inline ge ge_from_uniform_bytes(const uint8_t in[64]) {
    fe r1 = fe_frombytes(in), r2 = fe_frombytes(in + 32);
    return ge_add(ge_elligator(r1), ge_elligator(r2));
}

static inline ge ge_basepoint() {
    static const uint8_t B[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                                  0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
    ge p;
    ge_decompress(p, B);
    return p;
}

// plain double-and-add, scalar given as 256-bit little-endian integer (any value)
static inline ge ge_scalarmul_bytes(const uint8_t s[32], const ge &p) {
    ge r = ge_identity();
    for (int i = 255; i >= 0; i--) {
        r = ge_dbl(r);
        if ((s[i >> 3] >> (i & 7)) & 1) r = ge_add(r, p);
    }
    return r;
}
static inline ge ge_scalarmul(const sc &s, const ge &p) {
    uint8_t b[32];
    sc_tobytes(b, s);
    return ge_scalarmul_bytes(b, p);
}

static inline bool ge_is_identity(const ge &p) {  // Ristretto identity test (compress(p) == 0^32 semantics)
    return ge_eq(p, ge_identity());
}

}  // namespace orc
