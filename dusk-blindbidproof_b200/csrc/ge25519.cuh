// Ristretto255 / twisted-Edwards points for sm_100a (extended coordinates, a = -1), on top of fe25519.cuh.
// Device counterpart of curve25519-dalek 1.2.3's EdwardsPoint / RistrettoPoint / CompressedRistretto as the
// reference's hot path uses them through bulletproofs.
//
// Formats in HBM:
//   ge       extended (X:Y:Z:T), 4 x 32 B = 128 B  (partial sums, bucket sums, anything that is added *to*)
//   niels    affine (y+x, y-x, 2d*x*y), 3 x 32 B = 96 B  (bases that are added *from*: 7-multiplication mixed add)
#pragma once
#include "fe25519.cuh"

namespace bbp {

struct ge {
    fe X, Y, Z, T;
};
struct niels {
    fe yplusx, yminusx, xy2d;
};

BBP_DEV ge ge_identity() { ge r; r.X = fe_zero(); r.Y = fe_one(); r.Z = fe_one(); r.T = fe_zero(); return r; }

// extended + extended (unified, complete on the prime-order subgroup and its cosets): 9M
BBP_DEV ge ge_add(const ge &p, const ge &q) {
    fe A = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
    fe B = fe_mul(fe_add(p.Y, p.X), fe_add(q.Y, q.X));
    fe C = fe_mul(fe_mul(p.T, q.T), fe_d2());
    fe D = fe_dbl(fe_mul(p.Z, q.Z));
    fe E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
    ge r;
    r.X = fe_mul(E, F); r.Y = fe_mul(G, H); r.Z = fe_mul(F, G); r.T = fe_mul(E, H);
    return r;
}

// extended + affine niels (sign = true subtracts): 7M
BBP_DEV ge ge_madd(const ge &p, const niels &q, bool neg) {
    fe qa = fe_select(q.yminusx, q.yplusx, neg);
    fe qb = fe_select(q.yplusx, q.yminusx, neg);
    fe A = fe_mul(fe_sub(p.Y, p.X), qa);
    fe B = fe_mul(fe_add(p.Y, p.X), qb);
    fe C = fe_mul(p.T, q.xy2d);
    fe D = fe_dbl(p.Z);
    fe E = fe_sub(B, A), H = fe_add(B, A);
    fe F = fe_select(fe_sub(D, C), fe_add(D, C), neg);
    fe G = fe_select(fe_add(D, C), fe_sub(D, C), neg);
    ge r;
    r.X = fe_mul(E, F); r.Y = fe_mul(G, H); r.Z = fe_mul(F, G); r.T = fe_mul(E, H);
    return r;
}

// niels (optionally negated) as an extended point with Z = 2:  x = (yplusx - yminusx)/2, y = (yplusx + yminusx)/2
// (X:Y:Z:T) = (2x : 2y : 2 : 2xy) is the same projective point; 2xy = xy2d / d is not available, so T = X*Y/Z needs
// one multiplication by 1/2 folded in: T = (2x)(2y)/2.  Costs 1M + a halving-free trick: use Z = 2 and T = 2xy where
// (2x)(2y) = 4xy = 2 * (2xy) -> T = (2x)(2y) * inv2.  inv2 = (p+1)/2.
BBP_DEV ge ge_from_niels(const niels &q, bool neg) {
    fe x2 = fe_sub(q.yplusx, q.yminusx);   // 2x
    fe y2 = fe_add(q.yplusx, q.yminusx);   // 2y
    fe inv2;                                // (p+1)/2 = 2^254 - 9
    inv2.v[0] = 0xfffffff7u; inv2.v[1] = 0xffffffffu; inv2.v[2] = 0xffffffffu; inv2.v[3] = 0xffffffffu;
    inv2.v[4] = 0xffffffffu; inv2.v[5] = 0xffffffffu; inv2.v[6] = 0xffffffffu; inv2.v[7] = 0x3fffffffu;
    ge r;
    r.X = fe_cneg(x2, neg);
    r.Y = y2;
    r.Z = fe_zero(); r.Z.v[0] = 2;
    r.T = fe_mul(fe_mul(r.X, y2), inv2);
    return r;
}

// dedicated doubling (dbl-2008-hwcd with a = -1): 4S + 4M
BBP_DEV ge ge_dbl(const ge &p) {
    fe A = fe_sq(p.X), B = fe_sq(p.Y);
    fe C = fe_dbl(fe_sq(p.Z));
    fe D = fe_neg(A);
    fe E = fe_sub(fe_sub(fe_sq(fe_add(p.X, p.Y)), A), B);
    fe G = fe_add(D, B), F = fe_sub(G, C), H = fe_sub(D, B);
    ge r;
    r.X = fe_mul(E, F); r.Y = fe_mul(G, H); r.Z = fe_mul(F, G); r.T = fe_mul(E, H);
    return r;
}

BBP_DEV ge ge_neg(const ge &p) { ge r; r.X = fe_neg(p.X); r.Y = p.Y; r.Z = p.Z; r.T = fe_neg(p.T); return r; }

// affine (Z = 1) extended point -> niels
BBP_DEV niels ge_affine_to_niels(const fe &x, const fe &y, const fe &t) {
    niels n;
    n.yplusx = fe_add(y, x);
    n.yminusx = fe_sub(y, x);
    n.xy2d = fe_mul(t, fe_d2());
    return n;
}
// general extended point -> niels, given zinv = 1/Z
BBP_DEV niels ge_to_niels(const ge &p, const fe &zinv) {
    fe x = fe_mul(p.X, zinv), y = fe_mul(p.Y, zinv);
    return ge_affine_to_niels(x, y, fe_mul(x, y));
}

// RistrettoPoint::compress (SURVEY.md Appendix A "Compress"); writes 8 little-endian words
BBP_DEV void ge_compress_words(uint32_t *out, const ge &p) {
    fe X = p.X, Y = p.Y;
    fe u1 = fe_mul(fe_add(p.Z, Y), fe_sub(p.Z, Y));
    fe u2 = fe_mul(X, Y);
    fe I;
    fe_sqrt_ratio_i(I, fe_one(), fe_mul(u1, fe_sq(u2)));
    fe d1 = fe_mul(I, u1), d2 = fe_mul(I, u2);
    fe z_inv = fe_mul(fe_mul(d1, d2), p.T);
    bool rot = fe_isneg(fe_mul(p.T, z_inv));
    fe i = fe_sqrt_m1();
    fe nx = fe_mul(Y, i), ny = fe_mul(X, i);
    fe d_inv = fe_select(d2, fe_mul(d1, fe_invsqrt_a_minus_d()), rot);
    X = fe_select(X, nx, rot);
    Y = fe_select(Y, ny, rot);
    Y = fe_cneg(Y, fe_isneg(fe_mul(X, z_inv)));
    fe s = fe_abs(fe_mul(d_inv, fe_sub(p.Z, Y)));
    fe_tobytes_words(out, s);
}

// CompressedRistretto::decompress (Appendix A "Decompress"); false = invalid encoding
BBP_DEV bool ge_decompress_words(ge &out, const uint32_t *in) {
    fe s = fe_frombytes_words(in);
    fe sc = fe_canon(s);
    bool canonical = (in[7] >> 31) == 0;
#pragma unroll
    for (int k = 0; k < 8; k++) canonical = canonical && (sc.v[k] == s.v[k]);
    bool ok = canonical && !(in[0] & 1u);
    fe one = fe_one();
    fe ss = fe_sq(s);
    fe u1 = fe_sub(one, ss), u2 = fe_add(one, ss);
    fe u2s = fe_sq(u2);
    fe v = fe_sub(fe_neg(fe_mul(fe_d(), fe_sq(u1))), u2s);
    fe I;
    bool sq = fe_sqrt_ratio_i(I, one, fe_mul(v, u2s));
    fe Dx = fe_mul(I, u2);
    fe Dy = fe_mul(fe_mul(I, Dx), v);
    fe x = fe_abs(fe_mul(fe_dbl(s), Dx));
    fe y = fe_mul(u1, Dy);
    fe t = fe_mul(x, y);
    ok = ok && sq && !fe_isneg(t) && !fe_iszero(y);
    out.X = x; out.Y = y; out.Z = one; out.T = t;
    return ok;
}

// Elligator map of one field element (Appendix A MAP)
BBP_DEV ge ge_elligator(const fe &r0) {
    fe one = fe_one();
    fe minus_one = fe_neg(one);
    fe d = fe_d();
    fe r = fe_mul(fe_sqrt_m1(), fe_sq(r0));
    fe u = fe_mul(fe_add(r, one), fe_one_minus_d_sq());
    fe v = fe_mul(fe_sub(minus_one, fe_mul(r, d)), fe_add(r, d));
    fe s;
    bool sq = fe_sqrt_ratio_i(s, u, v);
    fe sp = fe_neg(fe_abs(fe_mul(s, r0)));
    s = fe_select(sp, s, sq);
    fe c = fe_select(r, minus_one, sq);
    fe N = fe_sub(fe_mul(fe_mul(c, fe_sub(r, one)), fe_d_minus_one_sq()), v);
    fe ss = fe_sq(s);
    fe w0 = fe_mul(fe_dbl(s), v);
    fe w1 = fe_mul(N, fe_sqrt_ad_minus_one());
    fe w2 = fe_sub(one, ss), w3 = fe_add(one, ss);
    ge p;
    p.X = fe_mul(w0, w3); p.Y = fe_mul(w2, w1); p.Z = fe_mul(w1, w3); p.T = fe_mul(w0, w2);
    return p;
}
// RistrettoPoint::from_uniform_bytes: 16 little-endian words
BBP_DEV ge ge_from_uniform_words(const uint32_t *in) {
    return ge_add(ge_elligator(fe_frombytes_words(in)), ge_elligator(fe_frombytes_words(in + 8)));
}

// Ristretto identity test (compress(p) == 0^32): X == 0 or Y == 0 on the encoding's coset representative
BBP_DEV bool ge_is_identity(const ge &p) { return fe_iszero(p.X) || fe_iszero(p.Y); }

BBP_DEV ge ge_basepoint() {
    const uint32_t bx[8] = FE_BASE_X_LIMBS, by[8] = FE_BASE_Y_LIMBS, bt[8] = FE_BASE_T_LIMBS;
    ge p;
    p.X = fe_const(bx); p.Y = fe_const(by); p.Z = fe_one(); p.T = fe_const(bt);
    return p;
}

// 128-bit vector load/store of 32-byte field elements (16-byte aligned addresses)
BBP_DEV fe fe_load(const void *ptr) {
    const uint4 *q = (const uint4 *)ptr;
    uint4 a = q[0], b = q[1];
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
BBP_DEV fe fe_load_ro(const void *ptr) {
    const uint4 *q = (const uint4 *)ptr;
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
BBP_DEV void fe_store(void *ptr, const fe &a) {
    uint4 *q = (uint4 *)ptr;
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
BBP_DEV ge ge_load(const void *ptr) {
    const char *c = (const char *)ptr;
    ge p;
    p.X = fe_load(c); p.Y = fe_load(c + 32); p.Z = fe_load(c + 64); p.T = fe_load(c + 96);
    return p;
}
BBP_DEV void ge_store(void *ptr, const ge &p) {
    char *c = (char *)ptr;
    fe_store(c, p.X); fe_store(c + 32, p.Y); fe_store(c + 64, p.Z); fe_store(c + 96, p.T);
}
BBP_DEV niels niels_load_ro(const void *ptr) {
    const char *c = (const char *)ptr;
    niels n;
    n.yplusx = fe_load_ro(c); n.yminusx = fe_load_ro(c + 32); n.xy2d = fe_load_ro(c + 64);
    return n;
}
BBP_DEV void niels_store(void *ptr, const niels &n) {
    char *c = (char *)ptr;
    fe_store(c, n.yplusx); fe_store(c + 32, n.yminusx); fe_store(c + 64, n.xy2d);
}

}  // namespace bbp
