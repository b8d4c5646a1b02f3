// ORACLE — test infrastructure only. Never linked into, imported by or called from the product
// (dusk-blindbidproof_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, and only as the checker / reported CPU baseline.
//
// GF(2^255-19) field arithmetic, radix 2^51 (5 x u64 limbs, unsigned __int128 products).
// Restates the serial u64 backend of curve25519-dalek 1.2.3 (Cargo.lock:141-143, un-vendored
// dependency of the reference). Parity is pinned in
// tests/test_oracle_primitives.py against Python big-int arithmetic, RFC 9496 vectors and libsodium.
#pragma once
#include <cstdint>
#include <cstring>

namespace orc {

typedef unsigned __int128 u128;

struct fe {
    uint64_t v[5];
};

static const uint64_t FE_MASK51 = (1ULL << 51) - 1;

static inline fe fe_zero() { return fe{{0, 0, 0, 0, 0}}; }
static inline fe fe_one() { return fe{{1, 0, 0, 0, 0}}; }

// carry chain bringing every limb below 2^51 (+ small excess in limb 0)
static inline fe fe_carry(fe a) {
    uint64_t c;
    c = a.v[0] >> 51; a.v[0] &= FE_MASK51; a.v[1] += c;
    c = a.v[1] >> 51; a.v[1] &= FE_MASK51; a.v[2] += c;
    c = a.v[2] >> 51; a.v[2] &= FE_MASK51; a.v[3] += c;
    c = a.v[3] >> 51; a.v[3] &= FE_MASK51; a.v[4] += c;
    c = a.v[4] >> 51; a.v[4] &= FE_MASK51; a.v[0] += 19 * c;
    c = a.v[0] >> 51; a.v[0] &= FE_MASK51; a.v[1] += c;
    return a;
}

static inline fe fe_add(const fe &a, const fe &b) {
    fe r;
    for (int i = 0; i < 5; i++) r.v[i] = a.v[i] + b.v[i];
    return fe_carry(r);
}

// a - b, computed as a + 16p - b so that no limb underflows (limbs of b are < 2^52)
static inline fe fe_sub(const fe &a, const fe &b) {
    fe r;
    r.v[0] = a.v[0] + 36028797018963664ULL - b.v[0];  // 16*(2^51-19)
    r.v[1] = a.v[1] + 36028797018963952ULL - b.v[1];  // 16*(2^51-1)
    r.v[2] = a.v[2] + 36028797018963952ULL - b.v[2];
    r.v[3] = a.v[3] + 36028797018963952ULL - b.v[3];
    r.v[4] = a.v[4] + 36028797018963952ULL - b.v[4];
    return fe_carry(r);
}

static inline fe fe_neg(const fe &a) { return fe_sub(fe_zero(), a); }

static inline fe fe_mul(const fe &a, const fe &b) {
    const uint64_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3], a4 = a.v[4];
    const uint64_t b0 = b.v[0], b1 = b.v[1], b2 = b.v[2], b3 = b.v[3], b4 = b.v[4];
    const uint64_t b1_19 = 19 * b1, b2_19 = 19 * b2, b3_19 = 19 * b3, b4_19 = 19 * b4;
    u128 c0 = (u128)a0 * b0 + (u128)a4 * b1_19 + (u128)a3 * b2_19 + (u128)a2 * b3_19 + (u128)a1 * b4_19;
    u128 c1 = (u128)a1 * b0 + (u128)a0 * b1 + (u128)a4 * b2_19 + (u128)a3 * b3_19 + (u128)a2 * b4_19;
    u128 c2 = (u128)a2 * b0 + (u128)a1 * b1 + (u128)a0 * b2 + (u128)a4 * b3_19 + (u128)a3 * b4_19;
    u128 c3 = (u128)a3 * b0 + (u128)a2 * b1 + (u128)a1 * b2 + (u128)a0 * b3 + (u128)a4 * b4_19;
    u128 c4 = (u128)a4 * b0 + (u128)a3 * b1 + (u128)a2 * b2 + (u128)a1 * b3 + (u128)a0 * b4;
    fe r;
    c1 += (uint64_t)(c0 >> 51); r.v[0] = (uint64_t)c0 & FE_MASK51;
    c2 += (uint64_t)(c1 >> 51); r.v[1] = (uint64_t)c1 & FE_MASK51;
    c3 += (uint64_t)(c2 >> 51); r.v[2] = (uint64_t)c2 & FE_MASK51;
    c4 += (uint64_t)(c3 >> 51); r.v[3] = (uint64_t)c3 & FE_MASK51;
    uint64_t carry = (uint64_t)(c4 >> 51); r.v[4] = (uint64_t)c4 & FE_MASK51;
    r.v[0] += carry * 19;
    r.v[1] += r.v[0] >> 51; r.v[0] &= FE_MASK51;
    return r;
}

static inline fe fe_sq(const fe &a) { return fe_mul(a, a); }

static inline fe fe_sqn(fe a, int n) {
    for (int i = 0; i < n; i++) a = fe_sq(a);
    return a;
}

// canonical little-endian encoding (fully reduced mod p)
static inline void fe_tobytes(uint8_t out[32], const fe &a) {
    fe t = fe_carry(a);
    // q = 1 iff t >= p : compute carry of t + 19 through all limbs
    uint64_t q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51;
    q = (t.v[2] + q) >> 51;
    q = (t.v[3] + q) >> 51;
    q = (t.v[4] + q) >> 51;
    t.v[0] += 19 * q;
    uint64_t c;
    c = t.v[0] >> 51; t.v[0] &= FE_MASK51; t.v[1] += c;
    c = t.v[1] >> 51; t.v[1] &= FE_MASK51; t.v[2] += c;
    c = t.v[2] >> 51; t.v[2] &= FE_MASK51; t.v[3] += c;
    c = t.v[3] >> 51; t.v[3] &= FE_MASK51; t.v[4] += c;
    t.v[4] &= FE_MASK51;
    uint64_t w0 = t.v[0] | (t.v[1] << 51);
    uint64_t w1 = (t.v[1] >> 13) | (t.v[2] << 38);
    uint64_t w2 = (t.v[2] >> 26) | (t.v[3] << 25);
    uint64_t w3 = (t.v[3] >> 39) | (t.v[4] << 12);
    memcpy(out, &w0, 8); memcpy(out + 8, &w1, 8); memcpy(out + 16, &w2, 8); memcpy(out + 24, &w3, 8);
}

// loads the low 255 bits (bit 255 ignored, as dalek's FieldElement::from_bytes does); value may be >= p
static inline fe fe_frombytes(const uint8_t in[32]) {
    uint64_t w0, w1, w2, w3;
    memcpy(&w0, in, 8); memcpy(&w1, in + 8, 8); memcpy(&w2, in + 16, 8); memcpy(&w3, in + 24, 8);
    fe r;
    r.v[0] = w0 & FE_MASK51;
    r.v[1] = ((w0 >> 51) | (w1 << 13)) & FE_MASK51;
    r.v[2] = ((w1 >> 38) | (w2 << 26)) & FE_MASK51;
    r.v[3] = ((w2 >> 25) | (w3 << 39)) & FE_MASK51;
    r.v[4] = (w3 >> 12) & FE_MASK51;
    return r;
}

static inline bool fe_eq(const fe &a, const fe &b) {
    uint8_t x[32], y[32];
    fe_tobytes(x, a); fe_tobytes(y, b);
    return memcmp(x, y, 32) == 0;
}
static inline bool fe_iszero(const fe &a) {
    uint8_t x[32];
    fe_tobytes(x, a);
    uint8_t acc = 0;
    for (int i = 0; i < 32; i++) acc |= x[i];
    return acc == 0;
}
static inline bool fe_isneg(const fe &a) {
    uint8_t x[32];
    fe_tobytes(x, a);
    return x[0] & 1;
}
static inline fe fe_cneg(const fe &a, bool neg) { return neg ? fe_neg(a) : a; }
static inline fe fe_abs(const fe &a) { return fe_cneg(a, fe_isneg(a)); }

// x^(2^250-1) and x^11, shared by invert and pow22523
static inline void fe_pow22501(fe &t19_out, fe &t3_out, const fe &x) {
    fe t0 = fe_sq(x);                  // 2
    fe t1 = fe_sqn(t0, 2);             // 8
    fe t2 = fe_mul(x, t1);             // 9
    fe t3 = fe_mul(t0, t2);            // 11
    fe t4 = fe_sq(t3);                 // 22
    fe t5 = fe_mul(t2, t4);            // 31 = 2^5-1
    fe t6 = fe_sqn(t5, 5);
    fe t7 = fe_mul(t6, t5);            // 2^10-1
    fe t8 = fe_sqn(t7, 10);
    fe t9 = fe_mul(t8, t7);            // 2^20-1
    fe t10 = fe_sqn(t9, 20);
    fe t11 = fe_mul(t10, t9);          // 2^40-1
    fe t12 = fe_sqn(t11, 10);
    fe t13 = fe_mul(t12, t7);          // 2^50-1
    fe t14 = fe_sqn(t13, 50);
    fe t15 = fe_mul(t14, t13);         // 2^100-1
    fe t16 = fe_sqn(t15, 100);
    fe t17 = fe_mul(t16, t15);         // 2^200-1
    fe t18 = fe_sqn(t17, 50);
    t19_out = fe_mul(t18, t13);        // 2^250-1
    t3_out = t3;
}

static inline fe fe_invert(const fe &x) {  // x^(p-2) = x^(2^255-21)
    fe t19, t3;
    fe_pow22501(t19, t3, x);
    fe t20 = fe_sqn(t19, 5);           // 2^255-32
    return fe_mul(t20, t3);            // 2^255-21
}

static inline fe fe_pow_p58(const fe &x) {  // x^((p-5)/8) = x^(2^252-3)
    fe t19, t3;
    fe_pow22501(t19, t3, x);
    fe t20 = fe_sqn(t19, 2);           // 2^252-4
    return fe_mul(x, t20);             // 2^252-3
}

static inline fe fe_from_u64(uint64_t x) {
    fe r = fe_zero();
    r.v[0] = x & FE_MASK51;
    r.v[1] = x >> 51;
    return r;
}

// curve / ristretto constants, derived numerically at start-up (no typed-in magic limbs) and
// cross-checked against the decimal values of SURVEY.md Appendix A in the tests.
struct fe_consts {
    fe d, d2, sqrt_m1, sqrt_ad_minus_one, invsqrt_a_minus_d, one_minus_d_sq, d_minus_one_sq;
};

static inline fe fe_pow_bytes(const fe &x, const uint8_t e[32]) {
    fe r = fe_one();
    for (int i = 255; i >= 0; i--) {
        r = fe_sq(r);
        if ((e[i >> 3] >> (i & 7)) & 1) r = fe_mul(r, x);
    }
    return r;
}

// sqrt_ratio_i (dalek field.rs FieldElement::sqrt_ratio_i)
static inline bool fe_sqrt_ratio_i(fe &out, const fe &u, const fe &v, const fe &sqrt_m1) {
    fe v3 = fe_mul(fe_sq(v), v);
    fe v7 = fe_mul(fe_sq(v3), v);
    fe r = fe_mul(fe_mul(u, v3), fe_pow_p58(fe_mul(u, v7)));
    fe check = fe_mul(v, fe_sq(r));
    fe neg_u = fe_neg(u);
    bool correct = fe_eq(check, u);
    bool flipped = fe_eq(check, neg_u);
    bool flipped_i = fe_eq(check, fe_mul(neg_u, sqrt_m1));
    if (flipped || flipped_i) r = fe_mul(r, sqrt_m1);
    out = fe_abs(r);
    return correct || flipped;
}

static inline fe_consts fe_make_constants() {
    fe_consts C;
    // d = -121665/121666
    C.d = fe_mul(fe_neg(fe_from_u64(121665)), fe_invert(fe_from_u64(121666)));
    C.d2 = fe_add(C.d, C.d);
    // sqrt(-1) = 2^((p-1)/4); (p-1)/4 = 2^253 - 5
    uint8_t e[32];
    memset(e, 0xff, 32);
    e[0] = 0xfb; e[31] = 0x1f;
    C.sqrt_m1 = fe_pow_bytes(fe_from_u64(2), e);
    fe one = fe_one();
    C.one_minus_d_sq = fe_sub(one, fe_sq(C.d));
    C.d_minus_one_sq = fe_sq(fe_sub(C.d, one));
    // sqrt(a*d - 1) with a = -1, i.e. sqrt(-d-1); dalek's constant is the odd ("negative") root
    fe adm1 = fe_sub(fe_neg(C.d), one);
    fe s;
    fe_sqrt_ratio_i(s, adm1, one, C.sqrt_m1);
    C.sqrt_ad_minus_one = fe_neg(s);
    // 1/sqrt(a-d) = 1/sqrt(-1-d), the even root
    fe t;
    fe_sqrt_ratio_i(t, one, adm1, C.sqrt_m1);
    C.invsqrt_a_minus_d = t;
    return C;
}

static inline const fe_consts &fe_constants() {
    static const fe_consts C = fe_make_constants();  // thread-safe (C++11 magic static)
    return C;
}

}  // namespace orc
