// ORACLE — test infrastructure only (see fe.h header). Scalars modulo the Ristretto group order
// l = 2^252 + 27742317777372353535851937790883648493.
// Restates curve25519-dalek 1.2.3 `Scalar` semantics used by the reference (SURVEY.md §2.2 U2,
// Appendix A "Scalars"): from_bytes_mod_order_wide (src/blindbid/mod.rs:16), from_bits
// (src/blindbid/bid.rs:27, verify.rs:115), from_canonical_bytes, + - * invert, batch_invert.
// Representation: 4 x u64 little-endian, always fully reduced (< l).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

typedef unsigned __int128 u128;

struct sc {
    uint64_t v[4];
};

static const uint64_t SC_L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0ULL, 0x1000000000000000ULL};

static inline sc sc_zero() { return sc{{0, 0, 0, 0}}; }
static inline sc sc_one() { return sc{{1, 0, 0, 0}}; }
static inline sc sc_from_u64(uint64_t x) { return sc{{x, 0, 0, 0}}; }

static inline bool sc_geq_l(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > SC_L[i]) return true;
        if (a[i] < SC_L[i]) return false;
    }
    return true;
}
static inline void sc_sub_l(uint64_t a[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - SC_L[i] - borrow;
        a[i] = (uint64_t)t;
        borrow = (t >> 64) & 1;
    }
}

static inline sc sc_add(const sc &a, const sc &b) {
    sc r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    // a,b < l < 2^253 so no overflow out of 256 bits
    if (sc_geq_l(r.v)) sc_sub_l(r.v);
    return r;
}
static inline sc sc_neg(const sc &a) {
    bool z = !(a.v[0] | a.v[1] | a.v[2] | a.v[3]);
    if (z) return a;
    sc r;
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)SC_L[i] - a.v[i] - borrow;
        r.v[i] = (uint64_t)t;
        borrow = (t >> 64) & 1;
    }
    return r;
}
static inline sc sc_sub(const sc &a, const sc &b) { return sc_add(a, sc_neg(b)); }

struct sc_mont_consts {
    uint64_t ninv;   // -l^{-1} mod 2^64
    sc r2;           // 2^512 mod l
    sc r1;           // 2^256 mod l
};

// Montgomery product a*b*2^-256 mod l (inputs < 2^256 with a*b < l*2^256, output < l)
static inline sc sc_montmul_raw(const uint64_t a[4], const uint64_t b[4], uint64_t ninv) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * ninv;
        c = (u128)m * SC_L[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * SC_L[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
        t[5] = 0;
    }
    sc r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || sc_geq_l(r.v)) sc_sub_l(r.v);
    return r;
}

static inline sc_mont_consts sc_make_consts() {
    sc_mont_consts C;
    // Newton iteration for l^{-1} mod 2^64
    uint64_t inv = 1;
    for (int i = 0; i < 7; i++) inv *= 2 - SC_L[0] * inv;
    C.ninv = (uint64_t)(0 - inv);
    // 2^k mod l by repeated doubling
    sc x = sc_one();
    for (int i = 0; i < 256; i++) x = sc_add(x, x);
    C.r1 = x;
    for (int i = 0; i < 256; i++) x = sc_add(x, x);
    C.r2 = x;
    return C;
}
static inline const sc_mont_consts &sc_consts() {
    static const sc_mont_consts C = sc_make_consts();
    return C;
}

static inline sc sc_mul(const sc &a, const sc &b) {
    const sc_mont_consts &C = sc_consts();
    sc t = sc_montmul_raw(a.v, b.v, C.ninv);      // a*b/R
    return sc_montmul_raw(t.v, C.r2.v, C.ninv);   // a*b
}

// reduce an arbitrary 256-bit little-endian integer
static inline sc sc_from_bytes_mod_order(const uint8_t in[32]) {
    const sc_mont_consts &C = sc_consts();
    uint64_t x[4];
    memcpy(x, in, 32);
    sc t = sc_montmul_raw(x, C.r2.v, C.ninv);     // x*R (x < 2^256, r2 < l)
    uint64_t one[4] = {1, 0, 0, 0};
    return sc_montmul_raw(t.v, one, C.ninv);      // x mod l
}
// dalek Scalar::from_bits: clear bit 255, keep the integer; the first arithmetic use reduces it.
// The oracle reduces immediately (equivalent: every dalek arithmetic output is canonical).
static inline sc sc_from_bits(const uint8_t in[32]) {
    uint8_t t[32];
    memcpy(t, in, 32);
    t[31] &= 0x7f;
    return sc_from_bytes_mod_order(t);
}
// 512-bit little-endian integer mod l (Scalar::from_bytes_mod_order_wide)
static inline sc sc_from_wide(const uint8_t in[64]) {
    const sc_mont_consts &C = sc_consts();
    sc lo = sc_from_bytes_mod_order(in);
    uint64_t hi[4];
    memcpy(hi, in + 32, 32);
    sc hr = sc_montmul_raw(hi, C.r2.v, C.ninv);   // hi * 2^256 mod l
    return sc_add(lo, hr);
}
static inline bool sc_from_canonical(sc &out, const uint8_t in[32]) {
    uint64_t x[4];
    memcpy(x, in, 32);
    if (sc_geq_l(x)) return false;
    memcpy(out.v, x, 32);
    return true;
}
static inline void sc_tobytes(uint8_t out[32], const sc &a) { memcpy(out, a.v, 32); }
static inline bool sc_eq(const sc &a, const sc &b) { return memcmp(a.v, b.v, 32) == 0; }
static inline bool sc_iszero(const sc &a) { return !(a.v[0] | a.v[1] | a.v[2] | a.v[3]); }

static inline sc sc_invert(const sc &a) {  // a^(l-2)
    uint64_t e[4] = {SC_L[0] - 2, SC_L[1], SC_L[2], SC_L[3]};
    sc r = sc_one();
    for (int i = 255; i >= 0; i--) {
        r = sc_mul(r, r);
        if ((e[i >> 6] >> (i & 63)) & 1) r = sc_mul(r, a);
    }
    return r;
}

// Scalar::batch_invert: inverts every entry in place, returns the product of all inverses
static inline sc sc_batch_invert(std::vector<sc> &xs) {
    size_t n = xs.size();
    std::vector<sc> pre(n);
    sc acc = sc_one();
    for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = sc_mul(acc, xs[i]); }
    sc inv = sc_invert(acc);
    sc ret = inv;
    for (size_t i = n; i-- > 0;) {
        sc t = sc_mul(inv, xs[i]);
        xs[i] = sc_mul(inv, pre[i]);
        inv = t;
    }
    return ret;
}

static inline sc sc_inner_product(const sc *a, const sc *b, size_t n) {
    sc acc = sc_zero();
    for (size_t i = 0; i < n; i++) acc = sc_add(acc, sc_mul(a[i], b[i]));
    return acc;
}

}  // namespace orc
