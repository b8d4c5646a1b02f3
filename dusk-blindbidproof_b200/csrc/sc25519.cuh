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
#include "fe25519.cuh"   // the 8 x 8 limb product (IMAD.WIDE carry chains) is shared with the field layer
#endif

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
BBP_HD sc sc_add(const sc &a, const sc &b) {
#ifdef __CUDA_ARCH__
    // branch-free: r = a + b (a, b < l < 2^253: no carry out), t = r - l, keep r when the subtraction borrows
    uint32_t r[8], t[8], bw;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(bw)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(sc_l_limb(0)), "r"(sc_l_limb(1)), "r"(sc_l_limb(2)), "r"(sc_l_limb(3)), "r"(sc_l_limb(4)), "r"(sc_l_limb(5)), "r"(sc_l_limb(6)), "r"(sc_l_limb(7)));
    sc o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.v[i] = bw ? r[i] : t[i];
    return o;
#else
    sc r;
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    if (sc_geq_l(r.v)) sc_sub_l(r.v);   // a, b < l < 2^253: no overflow out of 256 bits
    return r;
#endif
}
BBP_HD sc sc_sub(const sc &a, const sc &b) {
#ifdef __CUDA_ARCH__
    // branch-free: r = a - b, then add l back when the subtraction borrowed (a, b < l)
    uint32_t r[8], bw;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    sc o;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=&r"(o.v[0]), "=&r"(o.v[1]), "=&r"(o.v[2]), "=&r"(o.v[3]), "=&r"(o.v[4]), "=&r"(o.v[5]), "=&r"(o.v[6]), "=&r"(o.v[7])
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(sc_l_limb(0) & bw), "r"(sc_l_limb(1) & bw), "r"(sc_l_limb(2) & bw), "r"(sc_l_limb(3) & bw), "r"(sc_l_limb(4) & bw), "r"(sc_l_limb(5) & bw),
          "r"(sc_l_limb(6) & bw), "r"(sc_l_limb(7) & bw));
    return o;
#else
    return sc_add(a, sc_neg(b));
#endif
}


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
#if defined(__CUDACC__)
// -l^-1 mod 2^256 and c = l - 2^252 (125 bits)
#define SC_NPRIME_LIMBS {0x12547e1bu, 0xd2b51da3u, 0xfdba84ffu, 0xb1a206f2u, 0xffa36beau, 0x14e75438u, 0x6fe91836u, 0x9db6c6f2u}
#define SC_C_LIMBS {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu}

// r[0..11] = a[0..7] * b[0..3]: the first four rows of fe_mul_wide's even / odd column scheme
__device__ __forceinline__ void sc_mul_8x4_wide(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t ev[12], od[11];
    mul4(ev, a[0], a[2], a[4], a[6], b[0]);
    mul4(od, a[1], a[3], a[5], a[7], b[0]);
    od[8] = mad4_cc(od, a[0], a[2], a[4], a[6], b[1]);
    mad4_top_fresh(ev + 2, a[1], a[3], a[5], a[7], b[1]);
    ev[10] = mad4_cc(ev + 2, a[0], a[2], a[4], a[6], b[2]);
    mad4_top_half(od + 2, a[1], a[3], a[5], a[7], b[2]);
    od[10] = mad4_cc(od + 2, a[0], a[2], a[4], a[6], b[3]);
    mad4_top_half(ev + 4, a[1], a[3], a[5], a[7], b[3]);
    r[0] = ev[0];
    asm("add.cc.u32 %0, %11, %22;\n\t"
        "addc.cc.u32 %1, %12, %23;\n\t"
        "addc.cc.u32 %2, %13, %24;\n\t"
        "addc.cc.u32 %3, %14, %25;\n\t"
        "addc.cc.u32 %4, %15, %26;\n\t"
        "addc.cc.u32 %5, %16, %27;\n\t"
        "addc.cc.u32 %6, %17, %28;\n\t"
        "addc.cc.u32 %7, %18, %29;\n\t"
        "addc.cc.u32 %8, %19, %30;\n\t"
        "addc.cc.u32 %9, %20, %31;\n\t"
        "addc.u32 %10, %21, %32;"
        : "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11])
        : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]),
          "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]));
}

// Montgomery product on the device, separated-operand form: T = a b (16 limbs, the field layer's product);
// M = T_lo * (-l^-1) mod 2^256; T + M l has zero low half, and with l = 2^252 + c
//   (T + M l) / 2^256 = T_hi + floor((M c + M 2^252) / 2^256) + [T_lo != 0].
// ~160 IMAD.WIDE against ~350 mixed IMAD / IADD3 instructions of the limb-serial CIOS loop it replaces.
__device__ __forceinline__ sc sc_montmul_dev(const uint32_t *a, const uint32_t *b) {
    const uint32_t np[8] = SC_NPRIME_LIMBS, cl[4] = SC_C_LIMBS;
    uint32_t T[16], W[16], P[12];
    fe_mul_wide(T, a, b);
    fe_mul_wide(W, T, np);                    // only W[0..7] = M is used; the dead high rows are dropped by ptxas
    sc_mul_8x4_wide(P, W, cl);
    // Q = M << 252: Q[7] = M0 << 28, Q[8 + j] = (M[j] >> 4) | (M[j + 1] << 28), Q[15] = M7 >> 4
    uint32_t q7 = W[0] << 28, Q[8];
#pragma unroll
    for (int j = 0; j < 7; j++) Q[j] = (W[j] >> 4) | (W[j + 1] << 28);
    Q[7] = W[7] >> 4;
    uint32_t lo_nonzero = (T[0] | T[1] | T[2] | T[3] | T[4] | T[5] | T[6] | T[7]) != 0 ? 1u : 0u;
    sc r;
    uint32_t dummy;
    // U_hi = P[8..11] + Q + carry(P[7] + q7);   R = T_hi + U_hi + lo_nonzero
    asm("add.cc.u32 %8, %9, %10;\n\t"          // carry out of limb 7 of P + Q
        "addc.cc.u32 %0, %11, %15;\n\t"
        "addc.cc.u32 %1, %12, %16;\n\t"
        "addc.cc.u32 %2, %13, %17;\n\t"
        "addc.cc.u32 %3, %14, %18;\n\t"
        "addc.cc.u32 %4, %19, 0;\n\t"
        "addc.cc.u32 %5, %20, 0;\n\t"
        "addc.cc.u32 %6, %21, 0;\n\t"
        "addc.u32 %7, %22, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(dummy)
        : "r"(P[7]), "r"(q7), "r"(P[8]), "r"(P[9]), "r"(P[10]), "r"(P[11]), "r"(Q[0]), "r"(Q[1]), "r"(Q[2]), "r"(Q[3]), "r"(Q[4]), "r"(Q[5]), "r"(Q[6]),
          "r"(Q[7]));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+&r"(r.v[0]), "+&r"(r.v[1]), "+&r"(r.v[2]), "+&r"(r.v[3]), "+&r"(r.v[4]), "+&r"(r.v[5]), "+&r"(r.v[6]), "+&r"(r.v[7])
        : "r"(lo_nonzero));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+&r"(r.v[0]), "+&r"(r.v[1]), "+&r"(r.v[2]), "+&r"(r.v[3]), "+&r"(r.v[4]), "+&r"(r.v[5]), "+&r"(r.v[6]), "+&r"(r.v[7])
        : "r"(T[8]), "r"(T[9]), "r"(T[10]), "r"(T[11]), "r"(T[12]), "r"(T[13]), "r"(T[14]), "r"(T[15]));
    if (sc_geq_l(r.v)) sc_sub_l(r.v);
    return r;
}
#endif

BBP_HD sc sc_montmul(const uint32_t *a, const uint32_t *b) {
#if !defined(__CUDA_ARCH__)
    return sc_montmul_host64(a, b);
#elif !defined(BBP_SC_CIOS)
    return sc_montmul_dev(a, b);
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
// the same from 16 little-endian words
BBP_HD sc sc_from_wide_words(const uint32_t *w) {
    sc lo = sc_reduce_words(w);
    sc r2 = sc_r2();
    sc hi = sc_montmul(w + 8, r2.v);
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
