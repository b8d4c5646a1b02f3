// GF(2^255-19) for sm_100a: 8 x 32-bit saturated limbs, products on the FMA pipe as IMAD.WIDE.U32 carry chains.
// Device counterpart of the field layer under curve25519-dalek 1.2.3's FieldElement, which is what every point
// operation of the reference's hot path bottoms out in (the reference reaches it through
// bulletproofs from src/blindbid/proof.rs:88 and src/blindbid/verify.rs:88).
//
// Representation: an `fe` is any integer in [0, 2^256) congruent to the value mod p ("lazy" reduction; 2^256 = 38).
// Only fe_tobytes / fe_iszero / fe_isneg / fe_eq produce or look at the canonical representative.
//
// Multiplication layout: the 8x8 limb products are issued as 64-bit `mad.lo.cc/madc.hi.cc` pairs, which ptxas
// fuses into IMAD.WIDE.U32(.X). Products a[j]*b[i] with i+j even accumulate into `ev`, those with i+j odd into `od`
// (weight 2^32 higher), so that every carry chain is 64-bit aligned and four products long.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "curve_consts.inc"

namespace bbp {

struct fe {
    uint32_t v[8];
};

#define BBP_DEV __device__ __forceinline__

BBP_DEV fe fe_zero() { fe r; r.v[0] = 0; r.v[1] = 0; r.v[2] = 0; r.v[3] = 0; r.v[4] = 0; r.v[5] = 0; r.v[6] = 0; r.v[7] = 0; return r; }
BBP_DEV fe fe_one() { fe r = fe_zero(); r.v[0] = 1; return r; }

// c[0..7] += {x0,x1,x2,x3} * y as four 64-bit columns with a rippling carry; returns the carry out of c[7]
BBP_DEV uint32_t mad4_cc(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    uint32_t cy;
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+&r"(c[0]), "+&r"(c[1]), "+&r"(c[2]), "+&r"(c[3]), "+&r"(c[4]), "+&r"(c[5]), "+&r"(c[6]), "+&r"(c[7]), "=&r"(cy)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
    return cy;
}
// same, where c[6], c[7] have not been written yet (treated as zero); no carry out is possible
BBP_DEV void mad4_top_fresh(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, 0;\n\t"
        "madc.hi.u32 %7, %11, %12, 0;"
        : "+&r"(c[0]), "+&r"(c[1]), "+&r"(c[2]), "+&r"(c[3]), "+&r"(c[4]), "+&r"(c[5]), "=&r"(c[6]), "=&r"(c[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}
// same, where c[7] has not been written yet and c[6] holds a previous carry-out
BBP_DEV void mad4_top_half(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, 0;"
        : "+&r"(c[0]), "+&r"(c[1]), "+&r"(c[2]), "+&r"(c[3]), "+&r"(c[4]), "+&r"(c[5]), "+&r"(c[6]), "=&r"(c[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}
// c[0..7] = {x0,x1,x2,x3} * y (four independent 32x32->64 products)
BBP_DEV void mul4(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    uint64_t p0 = (uint64_t)x0 * y, p1 = (uint64_t)x1 * y, p2 = (uint64_t)x2 * y, p3 = (uint64_t)x3 * y;
    c[0] = (uint32_t)p0; c[1] = (uint32_t)(p0 >> 32);
    c[2] = (uint32_t)p1; c[3] = (uint32_t)(p1 >> 32);
    c[4] = (uint32_t)p2; c[5] = (uint32_t)(p2 >> 32);
    c[6] = (uint32_t)p3; c[7] = (uint32_t)(p3 >> 32);
}

// 512-bit product r[0..15] = a * b
BBP_DEV void fe_mul_wide(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t ev[16], od[15];
    // row 0: first touch of ev[0..7], od[0..7]
    mul4(ev, a[0], a[2], a[4], a[6], b[0]);
    mul4(od, a[1], a[3], a[5], a[7], b[0]);
    // row 1 (odd): a_even*b1 -> od[0..7] (+carry -> od[8]); a_odd*b1 -> ev[2..9] (ev[8], ev[9] fresh)
    od[8] = mad4_cc(od, a[0], a[2], a[4], a[6], b[1]);
    mad4_top_fresh(ev + 2, a[1], a[3], a[5], a[7], b[1]);
    // rows 2..7
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        // even row i: a_even*b[i] -> ev[i..i+7] (carry -> ev[i+8], fresh); a_odd*b[i] -> od[i..i+7] (od[i+7] fresh, od[i+6] carry)
        ev[i + 8] = mad4_cc(ev + i, a[0], a[2], a[4], a[6], b[i]);
        mad4_top_half(od + i, a[1], a[3], a[5], a[7], b[i]);
        // odd row i+1: a_even*b[i+1] -> od[i..i+7] (carry -> od[i+8], fresh); a_odd*b[i+1] -> ev[i+2..i+9] (ev[i+9] fresh, ev[i+8] carry)
        uint32_t cy = mad4_cc(od + i, a[0], a[2], a[4], a[6], b[i + 1]);
        if (i + 8 < 15) od[i + 8] = cy;   // the last row cannot carry out of the 512-bit product
        mad4_top_half(ev + i + 2, a[1], a[3], a[5], a[7], b[i + 1]);
    }
    // r = ev + (od << 32)
    r[0] = ev[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;"
        : "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]),
          "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
        : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]), "r"(ev[9]), "r"(ev[10]),
          "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]),
          "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]), "r"(od[9]),
          "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
}

// ---- dedicated squaring: 28 off-diagonal products (doubled by one funnel-shift pass) + 8 diagonal ones = 36 IMAD.WIDE against
// the 64 of fe_mul_wide. Chains of 1..3 products over the same even / odd column accumulators as the multiplication; the
// suffix says what the top of the chain meets: _cc = every slot already written, carry out returned; _top_fresh = the last
// product's two slots are unwritten; _top_half = its low slot holds an earlier carry-out, its high slot is unwritten.
BBP_DEV void mul3(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t y) {
    uint64_t p0 = (uint64_t)x0 * y, p1 = (uint64_t)x1 * y, p2 = (uint64_t)x2 * y;
    c[0] = (uint32_t)p0; c[1] = (uint32_t)(p0 >> 32);
    c[2] = (uint32_t)p1; c[3] = (uint32_t)(p1 >> 32);
    c[4] = (uint32_t)p2; c[5] = (uint32_t)(p2 >> 32);
}
BBP_DEV uint32_t mad3_cc(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t y) {
    uint32_t cy;
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, 0, 0;"
        : "+&r"(c[0]), "+&r"(c[1]), "+&r"(c[2]), "+&r"(c[3]), "+&r"(c[4]), "+&r"(c[5]), "=&r"(cy)
        : "r"(x0), "r"(x1), "r"(x2), "r"(y));
    return cy;
}
BBP_DEV uint32_t mad2_cc(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t y) {
    uint32_t cy;
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+&r"(c[0]), "+&r"(c[1]), "+&r"(c[2]), "+&r"(c[3]), "=&r"(cy)
        : "r"(x0), "r"(x1), "r"(y));
    return cy;
}
BBP_DEV uint32_t mad1_cc(uint32_t *c, uint32_t x0, uint32_t y) {
    uint32_t cy;
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, 0, 0;"
        : "+&r"(c[0]), "+&r"(c[1]), "=&r"(cy)
        : "r"(x0), "r"(y));
    return cy;
}
BBP_DEV void mad3_top_fresh(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %6, %9, %0;\n\t"
        "madc.hi.cc.u32 %1, %6, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %7, %9, %2;\n\t"
        "madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
        "madc.lo.cc.u32 %4, %8, %9, 0;\n\t"
        "madc.hi.u32 %5, %8, %9, 0;"
        : "+&r"(c[0]), "+&r"(c[1]), "+&r"(c[2]), "+&r"(c[3]), "=&r"(c[4]), "=&r"(c[5])
        : "r"(x0), "r"(x1), "r"(x2), "r"(y));
}
BBP_DEV void mad3_top_half(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %6, %9, %0;\n\t"
        "madc.hi.cc.u32 %1, %6, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %7, %9, %2;\n\t"
        "madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
        "madc.lo.cc.u32 %4, %8, %9, %4;\n\t"
        "madc.hi.u32 %5, %8, %9, 0;"
        : "+&r"(c[0]), "+&r"(c[1]), "+&r"(c[2]), "+&r"(c[3]), "+&r"(c[4]), "=&r"(c[5])
        : "r"(x0), "r"(x1), "r"(x2), "r"(y));
}
BBP_DEV void mad2_top_half(uint32_t *c, uint32_t x0, uint32_t x1, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"
        "madc.hi.u32 %3, %5, %6, 0;"
        : "+&r"(c[0]), "+&r"(c[1]), "+&r"(c[2]), "=&r"(c[3])
        : "r"(x0), "r"(x1), "r"(y));
}
BBP_DEV void mad1_top_half(uint32_t *c, uint32_t x0, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
        "madc.hi.u32 %1, %2, %3, 0;"
        : "+&r"(c[0]), "=&r"(c[1])
        : "r"(x0), "r"(y));
}

// 512-bit square r[0..15] = a * a. ev[k] is the limb of column k, od[k] the limb of column k + 1 (off-diagonal sums only).
BBP_DEV void fe_sq_wide(uint32_t *r, const uint32_t *a) {
    uint32_t ev[14], od[14];
    mul4(od, a[1], a[3], a[5], a[7], a[0]);                   // columns 1, 3, 5, 7
    mul3(ev + 2, a[2], a[4], a[6], a[0]);                     // columns 2, 4, 6
    od[8] = mad3_cc(od + 2, a[2], a[4], a[6], a[1]);          // columns 3, 5, 7
    mad3_top_fresh(ev + 4, a[3], a[5], a[7], a[1]);           // columns 4, 6, 8
    mad3_top_half(od + 4, a[3], a[5], a[7], a[2]);            // columns 5, 7, 9
    ev[10] = mad2_cc(ev + 6, a[4], a[6], a[2]);               // columns 6, 8
    od[10] = mad2_cc(od + 6, a[4], a[6], a[3]);               // columns 7, 9
    mad2_top_half(ev + 8, a[5], a[7], a[3]);                  // columns 8, 10
    mad2_top_half(od + 8, a[5], a[7], a[4]);                  // columns 9, 11
    ev[12] = mad1_cc(ev + 10, a[6], a[4]);                    // column 10
    od[12] = mad1_cc(od + 10, a[6], a[5]);                    // column 11
    mad1_top_half(ev + 12, a[7], a[5]);                       // column 12
    mad1_top_half(od + 12, a[7], a[6]);                       // column 13
    // s = ev + (od << 32): limbs 1 .. 14 (limb 0 and limb 15 of the off-diagonal sum are zero)
    uint32_t s[15];
    s[1] = od[0];
    asm("add.cc.u32 %0, %13, %25;\n\t"
        "addc.cc.u32 %1, %14, %26;\n\t"
        "addc.cc.u32 %2, %15, %27;\n\t"
        "addc.cc.u32 %3, %16, %28;\n\t"
        "addc.cc.u32 %4, %17, %29;\n\t"
        "addc.cc.u32 %5, %18, %30;\n\t"
        "addc.cc.u32 %6, %19, %31;\n\t"
        "addc.cc.u32 %7, %20, %32;\n\t"
        "addc.cc.u32 %8, %21, %33;\n\t"
        "addc.cc.u32 %9, %22, %34;\n\t"
        "addc.cc.u32 %10, %23, %35;\n\t"
        "addc.cc.u32 %11, %24, %36;\n\t"
        "addc.u32 %12, %37, 0;"
        : "=&r"(s[2]), "=&r"(s[3]), "=&r"(s[4]), "=&r"(s[5]), "=&r"(s[6]), "=&r"(s[7]), "=&r"(s[8]), "=&r"(s[9]), "=&r"(s[10]), "=&r"(s[11]), "=&r"(s[12]),
          "=&r"(s[13]), "=&r"(s[14])
        : "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]),
          "r"(ev[13]),
          "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]),
          "r"(od[12]), "r"(od[13]));
    // t = 2 s by independent funnel shifts (no carry chain), then r = t + sum_i a_i^2 2^(64 i): the diagonal products ride
    // the final addition as eight multiply-add pairs
    uint32_t t[16];
    t[1] = s[1] << 1;
#pragma unroll
    for (int k = 2; k < 15; k++) t[k] = __funnelshift_l(s[k - 1], s[k], 1);
    t[15] = s[14] >> 31;
    asm("mad.lo.cc.u32 %0, %16, %16, 0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %24;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %25;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %26;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %27;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %28;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %29;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %30;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %31;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %32;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %33;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %34;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %35;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %36;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %37;\n\t"
        "madc.hi.u32 %15, %23, %23, %38;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]),
          "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]),
          "r"(t[13]), "r"(t[14]), "r"(t[15]));
}

// 512-bit value -> fe: lo + 38*hi, then fold the small overflow twice.
// The fold is eight 32x32+64 products by 38. BBP_FOLD_SHIFT builds the alternative 38 hi = (hi << 5) + (hi << 2) + (hi << 1)
// (three funnel-shifted copies and three carry chains on the ALU pipe, 8 of the 72 wide products of a multiplication off the
// FMA-heavy pipe, which ncu shows 84 % busy in the bucket accumulation against 42 % for the ALU pipe). Measured on the 2^20
// MSM it is SLOWER — k_accumulate 1.10 -> 1.23 ms: 54 dependent ALU instructions per multiplication cost more issue slots
// and latency than the 8 multiplier slots they free — so the multiplier form stays the default.
BBP_DEV fe fe_reduce_wide(const uint32_t *r) {
    fe o;
    uint32_t lo[8], top;
#ifndef BBP_FOLD_SHIFT
    uint32_t od[8];
#pragma unroll
    for (int i = 0; i < 8; i++) lo[i] = r[i];
    top = mad4_cc(lo, r[8], r[10], r[12], r[14], 38u);
    mul4(od, r[9], r[11], r[13], r[15], 38u);
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+&r"(lo[1]), "+&r"(lo[2]), "+&r"(lo[3]), "+&r"(lo[4]), "+&r"(lo[5]), "+&r"(lo[6]), "+&r"(lo[7]), "+&r"(top)
        : "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
#else
    uint32_t s[9];
    // lo + (hi << 1)
#pragma unroll
    for (int i = 1; i < 8; i++) s[i] = __funnelshift_l(r[7 + i], r[8 + i], 1);
    s[0] = r[8] << 1; s[8] = r[15] >> 31;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, %25, 0;"
        : "=&r"(lo[0]), "=&r"(lo[1]), "=&r"(lo[2]), "=&r"(lo[3]), "=&r"(lo[4]), "=&r"(lo[5]), "=&r"(lo[6]), "=&r"(lo[7]), "=&r"(top)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]), "r"(s[8]));
    // + (hi << 2), + (hi << 5)
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        const int k = pass ? 5 : 2;
#pragma unroll
        for (int i = 1; i < 8; i++) s[i] = __funnelshift_l(r[7 + i], r[8 + i], k);
        s[0] = r[8] << k; s[8] = r[15] >> (32 - k);
        asm("add.cc.u32 %0, %0, %9;\n\t"
            "addc.cc.u32 %1, %1, %10;\n\t"
            "addc.cc.u32 %2, %2, %11;\n\t"
            "addc.cc.u32 %3, %3, %12;\n\t"
            "addc.cc.u32 %4, %4, %13;\n\t"
            "addc.cc.u32 %5, %5, %14;\n\t"
            "addc.cc.u32 %6, %6, %15;\n\t"
            "addc.cc.u32 %7, %7, %16;\n\t"
            "addc.u32 %8, %8, %17;"
            : "+&r"(lo[0]), "+&r"(lo[1]), "+&r"(lo[2]), "+&r"(lo[3]), "+&r"(lo[4]), "+&r"(lo[5]), "+&r"(lo[6]), "+&r"(lo[7]), "+&r"(top)
            : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]), "r"(s[8]));
    }
#endif
    // top <= 39 (multiplier form) / <= 38 (shift form); lo += 38*top
    uint32_t t = top * 38u, c2;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+&r"(lo[0]), "+&r"(lo[1]), "+&r"(lo[2]), "+&r"(lo[3]), "+&r"(lo[4]), "+&r"(lo[5]), "+&r"(lo[6]), "+&r"(lo[7]), "=&r"(c2)
        : "r"(t));
    lo[0] += (0u - c2) & 38u;   // after a wrap the value is < 1482, so this cannot carry
#pragma unroll
    for (int i = 0; i < 8; i++) o.v[i] = lo[i];
    return o;
}

BBP_DEV fe fe_mul(const fe &a, const fe &b) {
    uint32_t r[16];
    fe_mul_wide(r, a.v, b.v);
    return fe_reduce_wide(r);
}
#ifndef BBP_SQ_VIA_MUL
BBP_DEV fe fe_sq(const fe &a) {
    uint32_t r[16];
    fe_sq_wide(r, a.v);
    return fe_reduce_wide(r);
}
#else
BBP_DEV fe fe_sq(const fe &a) { return fe_mul(a, a); }
#endif

BBP_DEV fe fe_add(const fe &a, const fe &b) {
    fe r;
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    uint32_t t = (0u - c) & 38u, c2;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+&r"(r.v[0]), "+&r"(r.v[1]), "+&r"(r.v[2]), "+&r"(r.v[3]), "+&r"(r.v[4]), "+&r"(r.v[5]), "+&r"(r.v[6]), "+&r"(r.v[7]), "=&r"(c2)
        : "r"(t));
    r.v[0] += (0u - c2) & 38u;
    return r;
}

BBP_DEV fe fe_sub(const fe &a, const fe &b) {
    fe r;
    uint32_t bw;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // bw = 0 or 0xffffffff; a - b + 2^256 = a - b + 38 (mod p), so take the 38 back off
    uint32_t t = bw & 38u, b2;
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32 %8, 0, 0;"
        : "+&r"(r.v[0]), "+&r"(r.v[1]), "+&r"(r.v[2]), "+&r"(r.v[3]), "+&r"(r.v[4]), "+&r"(r.v[5]), "+&r"(r.v[6]), "+&r"(r.v[7]), "=&r"(b2)
        : "r"(t));
    r.v[0] -= b2 & 38u;   // after a second wrap the value is >= 2^256 - 38, so this cannot borrow
    return r;
}

BBP_DEV fe fe_neg(const fe &a) { return fe_sub(fe_zero(), a); }
BBP_DEV fe fe_dbl(const fe &a) { return fe_add(a, a); }

// branch-free select: c ? b : a
BBP_DEV fe fe_select(const fe &a, const fe &b, bool c) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c ? b.v[i] : a.v[i];
    return r;
}

// canonical representative in [0, p)
BBP_DEV fe fe_canon(const fe &a) {
    fe r = a;
    // fold bit 255: v = (v mod 2^255) + 19 * (v >> 255)  (< 2^255 + 19)
    uint32_t top = r.v[7] >> 31;
    r.v[7] &= 0x7fffffffu;
    uint32_t t = top * 19u;
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+&r"(r.v[0]), "+&r"(r.v[1]), "+&r"(r.v[2]), "+&r"(r.v[3]), "+&r"(r.v[4]), "+&r"(r.v[5]), "+&r"(r.v[6]), "+&r"(r.v[7])
        : "r"(t));
    // q = 1 iff v >= p, i.e. iff v + 19 >= 2^255
    fe s;
    asm("add.cc.u32 %0, %8, 19;\n\t"
        "addc.cc.u32 %1, %9, 0;\n\t"
        "addc.cc.u32 %2, %10, 0;\n\t"
        "addc.cc.u32 %3, %11, 0;\n\t"
        "addc.cc.u32 %4, %12, 0;\n\t"
        "addc.cc.u32 %5, %13, 0;\n\t"
        "addc.cc.u32 %6, %14, 0;\n\t"
        "addc.u32 %7, %15, 0;"
        : "=&r"(s.v[0]), "=&r"(s.v[1]), "=&r"(s.v[2]), "=&r"(s.v[3]), "=&r"(s.v[4]), "=&r"(s.v[5]), "=&r"(s.v[6]), "=&r"(s.v[7])
        : "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]));
    bool ge_p = (s.v[7] >> 31) != 0;
    s.v[7] &= 0x7fffffffu;
    return fe_select(r, s, ge_p);
}

BBP_DEV bool fe_iszero(const fe &a) {
    fe c = fe_canon(a);
    return (c.v[0] | c.v[1] | c.v[2] | c.v[3] | c.v[4] | c.v[5] | c.v[6] | c.v[7]) == 0;
}
BBP_DEV bool fe_isneg(const fe &a) { return fe_canon(a).v[0] & 1u; }
BBP_DEV bool fe_eq(const fe &a, const fe &b) { return fe_iszero(fe_sub(a, b)); }
BBP_DEV fe fe_cneg(const fe &a, bool neg) { return fe_select(a, fe_neg(a), neg); }
BBP_DEV fe fe_abs(const fe &a) { return fe_cneg(a, fe_isneg(a)); }

BBP_DEV fe fe_sqn(fe a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) a = fe_sq(a);
    return a;
}

// x^(2^250-1) and x^11
BBP_DEV void fe_pow22501(fe &t250, fe &t11, const fe &x) {
    fe t0 = fe_sq(x);
    fe t1 = fe_sqn(t0, 2);
    fe t2 = fe_mul(x, t1);       // 9
    fe t3 = fe_mul(t0, t2);      // 11
    fe t4 = fe_sq(t3);           // 22
    fe t5 = fe_mul(t2, t4);      // 2^5-1
    fe t6 = fe_mul(fe_sqn(t5, 5), t5);     // 2^10-1
    fe t7 = fe_mul(fe_sqn(t6, 10), t6);    // 2^20-1
    fe t8 = fe_mul(fe_sqn(t7, 20), t7);    // 2^40-1
    fe t9 = fe_mul(fe_sqn(t8, 10), t6);    // 2^50-1
    fe t10 = fe_mul(fe_sqn(t9, 50), t9);   // 2^100-1
    fe t11_ = fe_mul(fe_sqn(t10, 100), t10);  // 2^200-1
    t250 = fe_mul(fe_sqn(t11_, 50), t9);   // 2^250-1
    t11 = t3;
}
BBP_DEV fe fe_invert(const fe &x) {
    fe a, b;
    fe_pow22501(a, b, x);
    return fe_mul(fe_sqn(a, 5), b);
}
BBP_DEV fe fe_pow_p58(const fe &x) {
    fe a, b;
    fe_pow22501(a, b, x);
    return fe_mul(fe_sqn(a, 2), x);
}

BBP_DEV fe fe_const(const uint32_t (&k)[8]) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = k[i];
    return r;
}
#define BBP_FE_CONST(name, limbs) \
    BBP_DEV fe name() { const uint32_t k[8] = limbs; return fe_const(k); }
BBP_FE_CONST(fe_d, FE_D_LIMBS)
BBP_FE_CONST(fe_d2, FE_D2_LIMBS)
BBP_FE_CONST(fe_sqrt_m1, FE_SQRT_M1_LIMBS)
BBP_FE_CONST(fe_sqrt_ad_minus_one, FE_SQRT_AD_MINUS_ONE_LIMBS)
BBP_FE_CONST(fe_invsqrt_a_minus_d, FE_INVSQRT_A_MINUS_D_LIMBS)
BBP_FE_CONST(fe_one_minus_d_sq, FE_ONE_MINUS_D_SQ_LIMBS)
BBP_FE_CONST(fe_d_minus_one_sq, FE_D_MINUS_ONE_SQ_LIMBS)

// sqrt_ratio_i(u, v): returns was_square, writes the non-negative root
BBP_DEV bool fe_sqrt_ratio_i(fe &out, const fe &u, const fe &v) {
    fe v3 = fe_mul(fe_sq(v), v);
    fe v7 = fe_mul(fe_sq(v3), v);
    fe r = fe_mul(fe_mul(u, v3), fe_pow_p58(fe_mul(u, v7)));
    fe check = fe_mul(v, fe_sq(r));
    fe i = fe_sqrt_m1();
    fe neg_u = fe_neg(u);
    bool correct = fe_eq(check, u);
    bool flipped = fe_eq(check, neg_u);
    bool flipped_i = fe_eq(check, fe_mul(neg_u, i));
    fe ri = fe_mul(r, i);
    r = fe_select(r, ri, flipped || flipped_i);
    out = fe_abs(r);
    return correct || flipped;
}

// 32-byte little-endian (de)serialisation through 32-bit words (callers guarantee 4-byte alignment)
BBP_DEV fe fe_frombytes_words(const uint32_t *w) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
    r.v[7] &= 0x7fffffffu;   // bit 255 ignored, as dalek's FieldElement::from_bytes does
    return r;
}
BBP_DEV void fe_tobytes_words(uint32_t *w, const fe &a) {
    fe c = fe_canon(a);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = c.v[i];
}

}  // namespace bbp
