// Scalars modulo the Ristretto group order l = 2^252 + 27742317777372353535851937790883648493, host + device.
// Product-side counterpart of curve25519-dalek 1.2.3's `Scalar` as used by the reference (src/blindbid/mod.rs:16,
// bid.rs:27, verify.rs:115, proof.rs:57-64) and by bulletproofs underneath it.
// 8 x 32-bit limbs, Montgomery multiplication (CIOS) with R = 2^256; values are kept in plain (non-Montgomery)
// form and fully reduced, so `sc` compares / serialises directly.
#pragma once
#include <cstdint>
#include <cstring>
#include "curve_consts.inc"

#ifdef __CUDACC__
#define BBP_HD __host__ __device__ __forceinline__
#else
#define BBP_HD inline
#endif

namespace bbp {

struct sc {
    uint32_t v[8];
};

BBP_HD uint32_t sc_l_limb(int i) {
    const uint32_t L[8] = SC_L_LIMBS;
    return L[i];
}
BBP_HD uint32_t sc_r2_limb(int i) {
    const uint32_t K[8] = SC_R2_LIMBS;
    return K[i];
}

BBP_HD sc sc_zero() { sc r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
BBP_HD sc sc_one() { sc r = sc_zero(); r.v[0] = 1; return r; }
BBP_HD sc sc_from_u64(uint64_t x) { sc r = sc_zero(); r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); return r; }
BBP_HD bool sc_iszero(const sc &a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.v[i]; return o == 0; }
BBP_HD bool sc_eq(const sc &a, const sc &b) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i]; return o == 0; }

// a >= l ?
BBP_HD bool sc_geq_l(const uint32_t *a) {
    for (int i = 7; i >= 0; i--) {
        uint32_t li = sc_l_limb(i);
        if (a[i] > li) return true;
        if (a[i] < li) return false;
    }
    return true;
}
BBP_HD void sc_sub_l(uint32_t *a) {
    uint64_t borrow = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a[i] - sc_l_limb(i) - borrow;
        a[i] = (uint32_t)t;
        borrow = (t >> 32) & 1;
    }
}
BBP_HD sc sc_add(const sc &a, const sc &b) {
    sc r;
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    if (sc_geq_l(r.v)) sc_sub_l(r.v);   // a, b < l < 2^253: no overflow out of 256 bits
    return r;
}
BBP_HD sc sc_neg(const sc &a) {
    if (sc_iszero(a)) return a;
    sc r;
    uint64_t borrow = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)sc_l_limb(i) - a.v[i] - borrow;
        r.v[i] = (uint32_t)t;
        borrow = (t >> 32) & 1;
    }
    return r;
}
BBP_HD sc sc_sub(const sc &a, const sc &b) { return sc_add(a, sc_neg(b)); }

// Montgomery product a*b/2^256 mod l; inputs < 2^256 with a*b < l*2^256; output < l
#if !defined(__CUDA_ARCH__)
// host build: the same CIOS recurrence on 4 x 64-bit limbs (unsigned __int128 products)
inline sc sc_montmul_host64(const uint32_t *a32, const uint32_t *b32) {
    typedef unsigned __int128 u128;
    uint64_t a[4], b[4], L[4], t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        a[i] = (uint64_t)a32[2 * i] | ((uint64_t)a32[2 * i + 1] << 32);
        b[i] = (uint64_t)b32[2 * i] | ((uint64_t)b32[2 * i + 1] << 32);
        L[i] = (uint64_t)sc_l_limb(2 * i) | ((uint64_t)sc_l_limb(2 * i + 1) << 32);
    }
    // -l^-1 mod 2^64 from the 32-bit constant by one Newton step
    uint64_t n0 = (uint64_t)SC_N0INV;          // n0 * l == -1 mod 2^32
    n0 = n0 * (2 + L[0] * n0);                 // x' = x (2 + l x) keeps x l == -1 and doubles the precision
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
        uint64_t m = t[0] * n0;
        c = (u128)m * L[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * L[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    sc r;
    for (int i = 0; i < 4; i++) { r.v[2 * i] = (uint32_t)t[i]; r.v[2 * i + 1] = (uint32_t)(t[i] >> 32); }
    if (t[4] || sc_geq_l(r.v)) sc_sub_l(r.v);
    return r;
}
#endif
BBP_HD sc sc_montmul(const uint32_t *a, const uint32_t *b) {
#if !defined(__CUDA_ARCH__)
    return sc_montmul_host64(a, b);
#else
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a[j] * b[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * SC_N0INV;
        c = (uint64_t)m * sc_l_limb(0) + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)m * sc_l_limb(j) + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
    sc r;
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    if (t[8] || sc_geq_l(r.v)) sc_sub_l(r.v);
    return r;
#endif
}
BBP_HD sc sc_r2() { sc r; for (int i = 0; i < 8; i++) r.v[i] = sc_r2_limb(i); return r; }
BBP_HD sc sc_mul(const sc &a, const sc &b) {
    sc t = sc_montmul(a.v, b.v);
    sc r2 = sc_r2();
    return sc_montmul(t.v, r2.v);
}
// Montgomery-domain helpers for long product chains: to_mont(a) = a*R, from_mont(a) = a/R
BBP_HD sc sc_to_mont(const sc &a) { sc r2 = sc_r2(); return sc_montmul(a.v, r2.v); }
BBP_HD sc sc_from_mont(const sc &a) { sc one = sc_one(); return sc_montmul(a.v, one.v); }

// any 256-bit little-endian integer -> reduced scalar
BBP_HD sc sc_reduce_words(const uint32_t *w) {
    sc r2 = sc_r2();
    sc t = sc_montmul(w, r2.v);       // x*R mod l (x < 2^256, R2 < l)
    return sc_from_mont(t);
}
BBP_HD sc sc_from_bytes_mod_order(const uint8_t *in) {
    uint32_t w[8];
    for (int i = 0; i < 8; i++) w[i] = (uint32_t)in[4 * i] | ((uint32_t)in[4 * i + 1] << 8) | ((uint32_t)in[4 * i + 2] << 16) | ((uint32_t)in[4 * i + 3] << 24);
    return sc_reduce_words(w);
}
// Scalar::from_bits: clear bit 255; dalek keeps the integer unreduced until first use, all uses reduce
BBP_HD sc sc_from_bits(const uint8_t *in) {
    uint8_t t[32];
    for (int i = 0; i < 32; i++) t[i] = in[i];
    t[31] &= 0x7f;
    return sc_from_bytes_mod_order(t);
}
// Scalar::from_bytes_mod_order_wide
BBP_HD sc sc_from_wide(const uint8_t *in) {
    sc lo = sc_from_bytes_mod_order(in);
    uint32_t w[8];
    for (int i = 0; i < 8; i++) w[i] = (uint32_t)in[32 + 4 * i] | ((uint32_t)in[33 + 4 * i] << 8) | ((uint32_t)in[34 + 4 * i] << 16) | ((uint32_t)in[35 + 4 * i] << 24);
    sc r2 = sc_r2();
    sc hi = sc_montmul(w, r2.v);      // hi * 2^256 mod l
    return sc_add(lo, hi);
}
BBP_HD bool sc_from_canonical(sc &out, const uint8_t *in) {
    uint32_t w[8];
    for (int i = 0; i < 8; i++) w[i] = (uint32_t)in[4 * i] | ((uint32_t)in[4 * i + 1] << 8) | ((uint32_t)in[4 * i + 2] << 16) | ((uint32_t)in[4 * i + 3] << 24);
    if (sc_geq_l(w)) return false;
    for (int i = 0; i < 8; i++) out.v[i] = w[i];
    return true;
}
BBP_HD void sc_tobytes(uint8_t *out, const sc &a) {
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)a.v[i]; out[4 * i + 1] = (uint8_t)(a.v[i] >> 8);
        out[4 * i + 2] = (uint8_t)(a.v[i] >> 16); out[4 * i + 3] = (uint8_t)(a.v[i] >> 24);
    }
}

// a^(l-2)
BBP_HD sc sc_invert(const sc &a) {
    sc am = sc_to_mont(a);
    sc r = sc_to_mont(sc_one());
    for (int i = 252; i >= 0; i--) {
        r = sc_montmul(r.v, r.v);
        uint32_t e = sc_l_limb(i >> 5);
        if (i < 32) e = sc_l_limb(0) - 2;   // exponent l - 2 differs from l only in the low limb
        if ((e >> (i & 31)) & 1) r = sc_montmul(r.v, am.v);
    }
    return sc_from_mont(r);
}

}  // namespace bbp
