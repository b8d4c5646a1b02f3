// Scalar-vector (mod l) kernels of the R1CS prover / verifier and the inner-product argument:
// z / y power tables, the flattened-constraint sparse product, the l(x) / r(x) polynomial build with the t_1..t_6 inner
// products, the IPP round (fold + cross inner products + product-form L / R scalars) and the verifier's g / h / s
// assembly. They restate on the device what bulletproofs 1.0.4 @ 4a05305 computes inside Prover::prove,
// InnerProductProof::create and Verifier::verify (call sites src/blindbid/proof.rs:88, src/blindbid/verify.rs:88;
// algorithm SURVEY.md §8 a-5 .. a-8). One CTA per proof; a batch of proofs is one launch.
//
// Number format: every vector that lives across kernels is kept in Montgomery form (x * 2^256 mod l) so that a product
// is ONE CIOS multiplication; inputs are converted on first load, MSM scalars / proof scalars on the last store.
//
// The IPP never folds the generator vectors. Round j needs L_j = <a_lo, G_hi> + <b_hi, H_lo> + c_L Q over the folded
// bases; writing the folded bases out in terms of the ORIGINAL generators gives L_j = sum_i (a[..] * sG[i]) G[i] + ...
// with sG[i] the running product of u_k^(+-1) selected by the bits of i. L_j and R_j are therefore plain MSMs over the
// resident generator table with scalars computed here — the same group elements as the reference's folded form, hence
// the same compressed bytes — and the ~63 % of CPU time the reference spends folding G / H disappears.
#pragma once
#include "sc25519.cuh"

namespace bbp {

#define BBP_SC_THREADS 256
#define BBP_MAX_LONG 8      // long CSR rows handled block-cooperatively
#define BBP_LONG_ROW 24     // a row counts as long above this many entries

// per-proof challenge block (normal form, 32 B each)
enum chal_slot : uint32_t {
    CH_Y = 0, CH_YINV, CH_Z, CH_X, CH_U, CH_W, CH_UJ, CH_UJINV,   // prover + verifier
    CH_R, CH_A, CH_B, CH_TX, CH_TXBL, CH_EBL, CH_RHO,             // verifier only
    CH_UJ0,                                                        // verifier: u_0 .. u_{lg n - 1}, then their inverses
    CH_N = CH_UJ0 + 64
};

struct sc_batch {
    uint32_t n_proofs, n1, q, m, n, lg_n, n_pub;
    uint32_t gcols;                            // generator columns per family in the table (capacity x parties >= n): slot
                                               // layout [B, B_blinding, G[0..gcols), H[0..gcols)], slot length 2 + 2 gcols
    const uint32_t *row_ptr, *entries;         // circuit template CSR (shared by the batch)
    const uint32_t *coef;                      // per entry: index into the proof's public value table of the term's coefficient
                                               // (0 = one). nullptr when every variable coefficient is +-1 (the blind-bid circuit)
    uint32_t skip_ypow;                        // verifiers need only z^j and y^-i
    uint32_t n_long, long_rows[BBP_MAX_LONG];  // the few CSR rows long enough to be summed by the whole block (see flatten_long_rows)
    const uint32_t *const_j, *const_idx;       // constant terms (verifier)
    uint32_t n_const;
    const sc *chal;                            // [n_proofs][CH_N]
    sc *zpow, *ypow, *yinvpow;                 // [n_proofs][q], [n_proofs][n], [n_proofs][n]   (Montgomery)
    // prover
    const sc *aL, *aR, *aO, *sL, *sR;          // [n_proofs][n1] normal form
    const sc *vbl;                             // [n_proofs][m]
    const sc *blind3;                          // [n_proofs][3]: i_blinding, o_blinding, s_blinding
    sc *poly;                                  // [n_proofs][6][n1]: l1, l2, l3, r0, r1, r3 (Montgomery)
    sc *tout;                                  // [n_proofs][8]: t1..t6, t2_blinding, spare (normal form)
    sc *a, *b, *sG, *sH;                       // [n_proofs][n] (Montgomery)
    sc *slots;                                 // MSM scalar slots, 2 + 2n scalars each (normal form)
    sc *ab_out;                                // [n_proofs][2] final a, b
    // verifier
    const sc *pub;                             // [n_proofs][n_pub] public value tables (normal form): table indices >= n_pub_shared
    const sc *pub_shared;                      // [n_pub_shared] the table's prefix that every proof shares (blind bid: 1, MiMC constants)
    uint32_t n_pub_shared;
    sc *dyn_out;                               // [n_proofs][dyn_stride]: first m entries = rho * wV[i] * r * x^2 (written here)
    uint32_t dyn_stride;
    uint32_t dyn_done;                         // 1: k_dyn_weights has already written dyn_out (k_verify_scalars leaves it alone)
    sc *stat;                                  // [n_proofs][2 + 2 gcols] rho-weighted static-base scalars (Montgomery)
    sc *stab;                                  // [n_proofs][n] scratch: the IPP verification vector s (Montgomery)
    // hybrid IPP: after the first rounds the folded bases ARE materialised once (n_f per family), later rounds work on them
    uint32_t fac_n;                            // length of the live part of sG / sH (n, or n_f after materialisation)
    uint32_t late;                             // 0: slots over [B, B_bl, G, H] of the generator table; 1: [F_G, F_H, B]
    uint32_t compact;                          // early rounds: 1 = slots hold only the non-zero half of the columns
                                               // ([B | n/2 G terms | n/2 H terms], addressed through a per-round column map)
    sc *mat;                                   // [n_proofs][2 n] compact scalars of the materialisation MSM
    uint32_t shard_g, shard_G;                 // sharded IPP: only generator columns i = shard_g (mod shard_G) get scalars (G = 0 / 1: all)
    // aggregated range proofs (bulletproofs RangeProof::prove_multiple / verify_multiple)
    const uint64_t *rp_values;                 // [n_proofs][rp_m]
    uint32_t rp_bits, rp_m;                    // n = rp_bits * rp_m
};

__device__ __forceinline__ sc sc_mont_one() { return sc_to_mont(sc_one()); }
#ifndef BBP_MM_INLINE
#define BBP_MM_ATTR __noinline__
#else
#define BBP_MM_ATTR __forceinline__
#endif
__device__ BBP_MM_ATTR sc mm(sc a, sc b) { return sc_montmul(a.v, b.v); }

// base^e in the Montgomery domain, e < 2^16
__device__ inline sc sc_pow_small_mont(const sc &baseM, uint32_t e) {
    if (e == 0) return sc_mont_one();
    sc r = baseM;
    for (int bit = 30 - __clz(e); bit >= 0; bit--) {
        r = mm(r, r);
        if ((e >> bit) & 1) r = mm(r, baseM);
    }
    return r;
}

// sum over the block; result valid in every thread. smem: BBP_SC_THREADS entries.
__device__ inline sc block_sum_sc(sc v, sc *smem) {
    const uint32_t t = threadIdx.x;
    smem[t] = v;
    __syncthreads();
    for (uint32_t s = BBP_SC_THREADS / 2; s >= 1; s >>= 1) {
        if (t < s) smem[t] = sc_add(smem[t], smem[t + s]);
        __syncthreads();
    }
    sc r = smem[0];
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------- power tables
// zpow[j] = z^(j+1), j < q ; ypow[i] = y^i (unless B.skip_ypow), yinvpow[i] = y^-i, i < n. Two-level product form: with
// i = hi * 64 + lo, base^i = lo_tab[lo] * hi_tab[hi], lo_tab[lo] = base^lo (64 entries), hi_tab[hi] = (base^64)^hi; both
// small tables are assembled from the squaring chain base^(2^k) by the bits of their index, so an element costs ONE
// product and all of them are independent (the first version walked a chunk per thread: one product per element plus
// eight per thread to reach the chunk start, 9 k products per proof against 5.6 k here, on a 20-product dependency chain).
#define BBP_POW_LO_BITS 6
#define BBP_POW_MAX_HI 1024   // sequences of up to 65536 elements
__global__ void __launch_bounds__(BBP_SC_THREADS) k_powers(sc_batch B) {
    __shared__ sc sq[3][16];                     // base^(2^k)
    __shared__ sc lo_tab[3][1 << BBP_POW_LO_BITS];
    extern __shared__ sc hi_tab[];               // [3][n_hi]
    const uint32_t p = blockIdx.x, t = threadIdx.x;
    const sc *ch = B.chal + (size_t)p * CH_N;
    const uint32_t len[3] = {B.q ? B.q + 1 : 0u, B.skip_ypow ? 0u : B.n, B.n};   // z needs exponents 1 .. q
    uint32_t n_hi_max = 0;
    for (int s = 0; s < 3; s++) n_hi_max = max(n_hi_max, (len[s] + 63) >> BBP_POW_LO_BITS);
    if ((t & 31) == 0 && t < 96) {   // one thread of three different warps: z, y, y^-1
        const uint32_t s = t >> 5;
        if (len[s]) {
            sc r = sc_to_mont(ch[s == 0 ? CH_Z : s == 1 ? CH_Y : CH_YINV]);
            for (uint32_t k = 0; k < 16; k++) { sq[s][k] = r; r = mm(r, r); }
        }
    }
    __syncthreads();
    // small tables: entry e of lo_tab = prod over the bits of e of sq[k]; entry e of hi_tab = the same with sq[6 + k]
    for (uint32_t w = t; w < 3 * ((1u << BBP_POW_LO_BITS) + n_hi_max); w += BBP_SC_THREADS) {
        const uint32_t s = w / ((1u << BBP_POW_LO_BITS) + n_hi_max), e0 = w % ((1u << BBP_POW_LO_BITS) + n_hi_max);
        if (!len[s]) continue;
        const bool is_hi = e0 >= (1u << BBP_POW_LO_BITS);
        const uint32_t e = is_hi ? e0 - (1u << BBP_POW_LO_BITS) : e0;
        sc acc = sc_mont_one();
        bool first = true;
        for (uint32_t k = 0; k < 10; k++)
            if ((e >> k) & 1) {
                const sc &f = sq[s][(is_hi ? BBP_POW_LO_BITS : 0) + k];
                acc = first ? f : mm(acc, f);
                first = false;
            }
        if (is_hi) hi_tab[s * n_hi_max + e] = acc; else lo_tab[s][e] = acc;
    }
    __syncthreads();
    sc *out[3] = {B.zpow + (size_t)p * B.q, B.ypow + (size_t)p * B.n, B.yinvpow + (size_t)p * B.n};
    for (int s = 0; s < 3; s++) {
        if (!len[s]) continue;
        const uint32_t first = s == 0 ? 1u : 0u;   // zpow[j] holds exponent j + 1
        for (uint32_t i = first + t; i < len[s]; i += BBP_SC_THREADS) {
            const uint32_t lo = i & ((1u << BBP_POW_LO_BITS) - 1), hi = i >> BBP_POW_LO_BITS;
            sc v = hi == 0 ? lo_tab[s][lo] : (lo == 0 ? hi_tab[s * n_hi_max + hi] : mm(lo_tab[s][lo], hi_tab[s * n_hi_max + hi]));
            out[s][i - first] = v;
        }
    }
}
inline size_t k_powers_smem(uint32_t q, uint32_t n) { return (size_t)3 * (((q + 1 > n ? q + 1 : n) + 63) >> BBP_POW_LO_BITS) * sizeof(sc); }

// entry ci of a proof's public value table: the shared prefix lives once, the rest per proof
__device__ __forceinline__ sc pub_value(const sc_batch &B, const sc *pub, uint32_t ci) {
    return ci < B.n_pub_shared ? B.pub_shared[ci] : pub[ci - B.n_pub_shared];
}
// one CSR entry: z^(j+1) times the term's coefficient (Montgomery form). Circuits recorded through the generic constraint
// system carry arbitrary coefficients (pub[coef[e]], normal form); the blind-bid template has none (B.coef == nullptr).
__device__ __forceinline__ sc flatten_term(const sc_batch &B, const sc *zpow, const sc *pub, uint32_t e, uint32_t v) {
    sc zp = zpow[v & 0x7fffffffu];
    if (B.coef) {
        const uint32_t ci = B.coef[e];
        if (ci) zp = mm(zp, sc_to_mont(pub_value(B, pub, ci)));
    }
    return zp;
}
// signed sum of z powers over one CSR row (Montgomery form)
__device__ __forceinline__ sc flatten_row(const sc_batch &B, const sc *zpow, uint32_t row, const sc *pub) {
    sc acc = sc_zero();
    for (uint32_t e = B.row_ptr[row]; e < B.row_ptr[row + 1]; e++) {
        uint32_t v = B.entries[e];
        const sc zp = flatten_term(B, zpow, pub, e, v);
        acc = (v >> 31) ? sc_sub(acc, zp) : sc_add(acc, zp);
    }
    return acc;
}

// A wire that feeds many constraints has a long row (the blind-bid circuit has one a_O wire in 820 constraints and one in
// 279, every other row has <= 19 entries); walked by one thread it would bound the whole block, so the rows listed in
// B.long_rows are summed by all threads first and looked up afterwards.
__device__ inline void flatten_long_rows(const sc_batch &B, const sc *zpow, sc *long_val, sc *smem, const sc *pub) {
    for (uint32_t k = 0; k < B.n_long; k++) {
        const uint32_t row = B.long_rows[k];
        sc acc = sc_zero();
        for (uint32_t e = B.row_ptr[row] + threadIdx.x; e < B.row_ptr[row + 1]; e += BBP_SC_THREADS) {
            uint32_t v = B.entries[e];
            const sc zp = flatten_term(B, zpow, pub, e, v);
            acc = (v >> 31) ? sc_sub(acc, zp) : sc_add(acc, zp);
        }
        acc = block_sum_sc(acc, smem);
        if (threadIdx.x == 0) long_val[k] = acc;
    }
    __syncthreads();
}
__device__ __forceinline__ sc flatten_row_l(const sc_batch &B, const sc *zpow, uint32_t row, const sc *long_val, const sc *pub) {
    for (uint32_t k = 0; k < B.n_long; k++)
        if (B.long_rows[k] == row) return long_val[k];
    return flatten_row(B, zpow, row, pub);
}

// ---------------------------------------------------------------- prover: commitment scalar slots
// slot layout = generator table order: [B, B_blinding, G[0..n), H[0..n)]. Three slots per proof:
//   A_I1 = i_bl B_bl + <a_L, G> + <a_R, H>;  A_O1 = o_bl B_bl + <a_O, G>;  S1 = s_bl B_bl + <s_L, G> + <s_R, H>
// launch 1 (which0 = 0, n_which = 2): A_I1, A_O1 interleaved per proof into slots [0, 2P); launch 2 (which0 = 2, n_which = 1):
// S1 into slots [2P, 3P) — S1 needs the s_L / s_R draws, which may still be running on the RNG stream during launch 1.
__global__ void __launch_bounds__(BBP_SC_THREADS) k_commit_slots(sc_batch B, uint32_t which0, uint32_t n_which, uint32_t slot0) {
    const uint32_t p = blockIdx.x / n_which, which = which0 + blockIdx.x % n_which;
    const uint32_t slot_len = 2 + 2 * B.gcols;
    sc *out = B.slots + (size_t)(slot0 + blockIdx.x) * slot_len;
    const sc *g = (which == 0 ? B.aL : which == 1 ? B.aO : B.sL) + (size_t)p * B.n1;
    const sc *h = (which == 0 ? B.aR : which == 1 ? nullptr : B.sR);
    if (h) h += (size_t)p * B.n1;
    for (uint32_t i = threadIdx.x; i < slot_len; i += BBP_SC_THREADS) {
        sc v = sc_zero();
        if (i == 1) v = B.blind3[(size_t)p * 3 + which];
        else if (i >= 2 && i < 2 + B.gcols) { if (i - 2 < B.n1) v = g[i - 2]; }
        else if (i >= 2 + B.gcols) { if (h && i - 2 - B.gcols < B.n1) v = h[i - 2 - B.gcols]; }
        out[i] = v;
    }
}

// ---------------------------------------------------------------- prover: witness of the blind-bid circuit
// One thread per proof evaluates the circuit of src/gadgets.rs natively — the same walk the host evaluator (circuit.h)
// does through the generic gadget code — and writes the multiplier rows (a_L, a_R, a_O) straight into the witness arrays:
//   MiMC(k, 0) -> m, MiMC(d, m) -> x, booleans t_i (1 - t_i), pairs (item_i, t_i), (t_i, x), MiMC(seed, x) -> y,
//   MiMC(seed, m), (y, y_inv), (d, y_inv)                       [src/gadgets.rs:6-34; multiplier order = generator index]
// in: [n_proofs][4 + L] = d, k, y_inv, seed, items[0..L); toggle[p] = index of the bidder's own item; mimc_c = 90 constants.
struct witness_rows {
    sc *aL, *aR, *aO;
    uint32_t i;
    __device__ __forceinline__ sc mul(const sc &l, const sc &r) {
        sc o = sc_mul(l, r);
        aL[i] = l; aR[i] = r; aO[i] = o;
        i++;
        return o;
    }
};
__device__ inline sc mimc_rows(witness_rows &W, const sc &left, const sc &key, const sc *__restrict__ c) {
    sc x = left;
    for (uint32_t r = 0; r < 90; r++) {
        sc a = sc_add(sc_add(x, key), c[r]);
        sc a2 = W.mul(a, a);
        sc a3 = W.mul(a2, a);
        sc a4 = W.mul(a2, a2);
        x = W.mul(a4, a3);
    }
    return sc_add(x, key);
}
__global__ void __launch_bounds__(32) k_blindbid_witness(const sc *__restrict__ in, const uint32_t *__restrict__ toggle, const sc *__restrict__ mimc_c,
                                                         uint32_t n_proofs, uint32_t L, uint32_t n1, sc *__restrict__ aL, sc *__restrict__ aR, sc *__restrict__ aO) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_proofs) return;
    const sc *v = in + (size_t)p * (4 + L);
    const sc d = v[0], k = v[1], y_inv = v[2], seed = v[3];
    witness_rows W = {aL + (size_t)p * n1, aR + (size_t)p * n1, aO + (size_t)p * n1, 0};
    sc m = mimc_rows(W, k, sc_zero(), mimc_c);
    sc x = mimc_rows(W, d, m, mimc_c);
    const uint32_t tg = toggle[p];
    for (uint32_t i = 0; i < L; i++) {               // boolean_gadget: t (1 - t)
        sc t = (i == tg) ? sc_one() : sc_zero();
        W.mul(t, sc_sub(sc_one(), t));
    }
    for (uint32_t i = 0; i < L; i++) {               // one_of_many: item_i * t_i, t_i * x
        sc t = (i == tg) ? sc_one() : sc_zero();
        W.mul(v[4 + i], t);
        W.mul(t, x);
    }
    sc y = mimc_rows(W, seed, x, mimc_c);
    mimc_rows(W, seed, m, mimc_c);
    W.mul(y, y_inv);
    W.mul(d, y_inv);
}

// ---------------------------------------------------------------- prover: l / r polynomials and t_1 .. t_6
__global__ void __launch_bounds__(BBP_SC_THREADS, 2) k_polys(sc_batch B) {
    __shared__ sc smem[BBP_SC_THREADS];
    const uint32_t p = blockIdx.x, t = threadIdx.x, n1 = B.n1;
    const sc *zpow = B.zpow + (size_t)p * B.q, *ypow = B.ypow + (size_t)p * B.n, *yinv = B.yinvpow + (size_t)p * B.n;
    sc *poly = B.poly + (size_t)p * 6 * n1;
    __shared__ sc long_val[BBP_MAX_LONG];
    const sc *pub = B.pub ? B.pub + (size_t)p * B.n_pub : nullptr;   // coefficient table (generic circuits only)
    flatten_long_rows(B, zpow, long_val, smem, pub);
    sc t1 = sc_zero(), t2 = t1, t3 = t1, t4 = t1, t5 = t1, t6 = t1;
    for (uint32_t i = t; i < n1; i += BBP_SC_THREADS) {
        sc wL = flatten_row_l(B, zpow, i, long_val, pub), wR = flatten_row_l(B, zpow, n1 + i, long_val, pub), wO = flatten_row_l(B, zpow, 2 * n1 + i, long_val, pub);
        size_t k = (size_t)p * n1 + i;
        sc aL = sc_to_mont(B.aL[k]), aR = sc_to_mont(B.aR[k]), aO = sc_to_mont(B.aO[k]), sL = sc_to_mont(B.sL[k]), sR = sc_to_mont(B.sR[k]);
        sc l1 = sc_add(aL, mm(yinv[i], wR));
        sc r0 = sc_sub(wO, ypow[i]);
        sc r1 = sc_add(mm(ypow[i], aR), wL);
        sc r3 = mm(ypow[i], sR);
        poly[i] = l1; poly[n1 + i] = aO; poly[2 * n1 + i] = sL; poly[3 * n1 + i] = r0; poly[4 * n1 + i] = r1; poly[5 * n1 + i] = r3;
        t1 = sc_add(t1, mm(l1, r0));
        t2 = sc_add(t2, sc_add(mm(l1, r1), mm(aO, r0)));
        t3 = sc_add(t3, sc_add(mm(aO, r1), mm(sL, r0)));
        t4 = sc_add(t4, sc_add(mm(l1, r3), mm(sL, r1)));
        t5 = sc_add(t5, mm(aO, r3));
        t6 = sc_add(t6, mm(sL, r3));
    }
    // t2_blinding = <wV, v_blinding>
    sc tb = sc_zero();
    for (uint32_t i = t; i < B.m; i += BBP_SC_THREADS) {
        sc wV = flatten_row(B, zpow, 3 * n1 + i, pub);
        tb = sc_add(tb, mm(wV, sc_to_mont(B.vbl[(size_t)p * B.m + i])));
    }
    sc r[7] = {t1, t2, t3, t4, t5, t6, tb};
    for (int k = 0; k < 7; k++) {
        sc s = block_sum_sc(r[k], smem);
        if (t == 0) B.tout[(size_t)p * 8 + k] = sc_from_mont(s);
    }
}

// ---------------------------------------------------------------- prover: IPP set-up
// a = l(x), b = r(x) padded to n (l = 0, r = -y^i beyond n1); sG = G_factors (1 | u), sH = y^-i * G_factors
__global__ void __launch_bounds__(BBP_SC_THREADS) k_ipp_init(sc_batch B) {
    const uint32_t p = blockIdx.x, n1 = B.n1;
    const sc *ch = B.chal + (size_t)p * CH_N;
    const sc *poly = B.poly + (size_t)p * 6 * n1, *ypow = B.ypow + (size_t)p * B.n, *yinv = B.yinvpow + (size_t)p * B.n;
    sc xM = sc_to_mont(ch[CH_X]), uM = sc_to_mont(ch[CH_U]);
    sc x2 = mm(xM, xM), x3 = mm(x2, xM), one = sc_mont_one();
    for (uint32_t i = threadIdx.x; i < B.n; i += BBP_SC_THREADS) {
        sc l, r, gf;
        if (i < n1) {
            l = sc_add(sc_add(mm(poly[i], xM), mm(poly[n1 + i], x2)), mm(poly[2 * n1 + i], x3));
            r = sc_add(sc_add(poly[3 * n1 + i], mm(poly[4 * n1 + i], xM)), mm(poly[5 * n1 + i], x3));
            gf = one;
        } else {
            l = sc_zero();
            r = sc_neg(ypow[i]);
            gf = uM;
        }
        size_t k = (size_t)p * B.n + i;
        B.a[k] = l; B.b[k] = r; B.sG[k] = gf; B.sH[k] = mm(yinv[i], gf);
    }
}

// ---------------------------------------------------------------- IPP round j (0-based)
// 1. for j > 0: fold a, b with the previous challenge (CH_UJ / CH_UJINV) and extend the product-form factors sG, sH
// 2. c_L = <a_lo, b_hi>, c_R = <a_hi, b_lo>
// 3. write the L slot (2p) and R slot (2p + 1): coefficient on B is c * w (Q = w B), B_blinding gets 0
// With fold_only the kernel stops after step 1 on the final length-1 vectors and emits a, b.
__global__ void __launch_bounds__(BBP_SC_THREADS) k_ipp_round(sc_batch B, uint32_t j, uint32_t mode) {
    __shared__ sc smem[BBP_SC_THREADS];
    const uint32_t p = blockIdx.x, t = threadIdx.x, n = B.n, fn = B.fac_n;
    const bool fold_only = mode & 1, skip_fold = mode & 2;
    const sc *ch = B.chal + (size_t)p * CH_N;
    sc *a = B.a + (size_t)p * n, *b = B.b + (size_t)p * n, *sG = B.sG + (size_t)p * n, *sH = B.sH + (size_t)p * n;
    const uint32_t nj = n >> j;           // current vector length
    if (j > 0 && !skip_fold) {
        sc uM = sc_to_mont(ch[CH_UJ]), uiM = sc_to_mont(ch[CH_UJINV]);
        for (uint32_t k = t; k < nj; k += BBP_SC_THREADS) {
            a[k] = sc_add(mm(a[k], uM), mm(uiM, a[nj + k]));
            b[k] = sc_add(mm(b[k], uiM), mm(uM, b[nj + k]));
        }
        if (!fold_only) {
            const uint32_t n_old = nj << 1;
            for (uint32_t i = t; i < fn; i += BBP_SC_THREADS) {
                bool lo = (i & (n_old - 1)) < nj;
                sG[i] = mm(sG[i], lo ? uiM : uM);
                sH[i] = mm(sH[i], lo ? uM : uiM);
            }
        }
        __syncthreads();
    }
    if (fold_only) {
        if (t == 0) { B.ab_out[(size_t)p * 2] = sc_from_mont(a[0]); B.ab_out[(size_t)p * 2 + 1] = sc_from_mont(b[0]); }
        return;
    }
    const uint32_t nh = nj >> 1;
    sc cl = sc_zero(), cr = sc_zero();
    for (uint32_t k = t; k < nh; k += BBP_SC_THREADS) {
        cl = sc_add(cl, mm(a[k], b[nh + k]));
        cr = sc_add(cr, mm(a[nh + k], b[k]));
    }
    cl = block_sum_sc(cl, smem);
    cr = block_sum_sc(cr, smem);
    if (B.compact && !B.late) {
        // compact early-round slots: L = [c_L w | a_lo * G_hi | b_hi * H_lo], R = [c_R w | a_hi * G_lo | b_lo * H_hi], n/2 terms each,
        // enumerated block by block (tt = block * nh + off); the engine maps entry e to its generator column (ipp_colmap)
        const uint32_t slot_len = 1 + n, half = n >> 1;
        sc *sl = B.slots + (size_t)(2 * p) * slot_len, *sr = sl + slot_len;
        if (t == 0) {
            sc wM = sc_to_mont(ch[CH_W]);
            sl[0] = sc_from_mont(mm(cl, wM)); sr[0] = sc_from_mont(mm(cr, wM));
        }
        for (uint32_t tt = t; tt < half; tt += BBP_SC_THREADS) {
            uint32_t blk = tt / nh, off = tt % nh, lo = blk * nj + off, hi = lo + nh;
            sl[1 + tt] = sc_from_mont(mm(a[off], sG[hi]));
            sl[1 + half + tt] = sc_from_mont(mm(b[off + nh], sH[lo]));
            sr[1 + tt] = sc_from_mont(mm(a[off + nh], sG[lo]));
            sr[1 + half + tt] = sc_from_mont(mm(b[off], sH[hi]));
        }
        return;
    }
    // slot layout: early rounds [B, B_bl, G[0..gc), H[0..gc)]; late rounds [F_G[0..fn), F_H[0..fn), B]
    const uint32_t gc = B.late ? fn : B.gcols, hdr = B.late ? 0 : 2, slot_len = B.late ? 2 * fn + 1 : 2 + 2 * gc;
    sc *sl = B.slots + (size_t)(2 * p) * slot_len, *sr = sl + slot_len;
    const bool sharded = B.shard_G > 1;
    if (t == 0) {
        sc wM = sc_to_mont(ch[CH_W]);
        uint32_t bpos = B.late ? 2 * fn : 0;
        const bool own_q = !sharded || B.shard_g == 0;   // the c * Q term belongs to shard 0
        sl[bpos] = own_q ? sc_from_mont(mm(cl, wM)) : sc_zero(); sr[bpos] = own_q ? sc_from_mont(mm(cr, wM)) : sc_zero();
        if (!B.late) { sl[1] = sc_zero(); sr[1] = sc_zero(); }
    }
    for (uint32_t i = t; i < gc; i += BBP_SC_THREADS) {
        sc zero = sc_zero();
        if (i >= fn || (sharded && i % B.shard_G != B.shard_g)) {    // columns beyond the vector length, or another shard's
            sl[hdr + i] = zero; sr[hdr + i] = zero; sl[hdr + gc + i] = zero; sr[hdr + gc + i] = zero;
            continue;
        }
        uint32_t k = i & (nj - 1);
        if (k >= nh) {   // base i sits in the high half of its block: L takes a_lo * G_hi, R takes b_lo * H_hi
            sl[hdr + i] = sc_from_mont(mm(a[k - nh], sG[i]));
            sr[hdr + i] = zero;
            sl[hdr + gc + i] = zero;
            sr[hdr + gc + i] = sc_from_mont(mm(b[k - nh], sH[i]));
        } else {          // low half: R takes a_hi * G_lo, L takes b_hi * H_lo
            sl[hdr + i] = zero;
            sr[hdr + i] = sc_from_mont(mm(a[k + nh], sG[i]));
            sl[hdr + gc + i] = sc_from_mont(mm(b[k + nh], sH[i]));
            sr[hdr + gc + i] = zero;
        }
    }
}

// standalone InnerProductProof::create (bbp_ipp_create): a, b and the G / H factors arrive in normal form
__global__ void __launch_bounds__(BBP_SC_THREADS) k_ipp_load(sc_batch B, const sc *__restrict__ a, const sc *__restrict__ b, const sc *__restrict__ gf,
                                                             const sc *__restrict__ hf) {
    const uint32_t p = blockIdx.x, n = B.n;
    for (uint32_t k = threadIdx.x; k < n; k += BBP_SC_THREADS) {
        const size_t o = (size_t)p * n + k;
        B.a[o] = sc_to_mont(a[o]); B.b[o] = sc_to_mont(b[o]); B.sG[o] = sc_to_mont(gf[o]); B.sH[o] = sc_to_mont(hf[o]);
    }
}

// Materialisation step of the hybrid IPP, run once when the vectors have shrunk to n_f = n >> j0: applies the pending
// fold (challenge of round j0 - 1), then emits the compact scalars of F_G[k] = sum_e sG[k + n_f e] G[k + n_f e] (and F_H):
// mat[fam][k][e], k < n_f, e < n / n_f — one MSM slot of n / n_f scalars per folded base — and resets the factors to 1.
__global__ void __launch_bounds__(BBP_SC_THREADS) k_ipp_materialize(sc_batch B, uint32_t j0) {
    const uint32_t p = blockIdx.x, t = threadIdx.x, n = B.n;
    const sc *ch = B.chal + (size_t)p * CH_N;
    sc *a = B.a + (size_t)p * n, *b = B.b + (size_t)p * n, *sG = B.sG + (size_t)p * n, *sH = B.sH + (size_t)p * n;
    const uint32_t nf = n >> j0, n_old = nf << 1, E = n / nf;
    sc uM = sc_to_mont(ch[CH_UJ]), uiM = sc_to_mont(ch[CH_UJINV]);
    for (uint32_t k = t; k < nf; k += BBP_SC_THREADS) {
        a[k] = sc_add(mm(a[k], uM), mm(uiM, a[nf + k]));
        b[k] = sc_add(mm(b[k], uiM), mm(uM, b[nf + k]));
    }
    sc *mat = B.mat + (size_t)p * 2 * n;
    for (uint32_t idx = t; idx < 2 * n; idx += BBP_SC_THREADS) {
        uint32_t fam = idx / n, r = idx % n, k = r / E, e = r % E, src = k + nf * e;
        bool lo = (src & (n_old - 1)) < nf;
        sc f = fam ? mm(sH[src], lo ? uM : uiM) : mm(sG[src], lo ? uiM : uM);
        mat[idx] = sc_from_mont(f);
    }
    __syncthreads();
    sc one = sc_mont_one();
    for (uint32_t k = t; k < nf; k += BBP_SC_THREADS) { sG[k] = one; sH[k] = one; }
}

// InnerProductProof::verification_scalars: s[0] = prod u_j^-1, s[i] = s[i - 2^b] * u_{lg-1-b}^2 for 2^b <= i < 2^(b+1):
// one multiplication per element, lg n block-wide steps (the whole CTA works on one proof). uj: u_0.., then inverses.
__device__ inline void build_s_table(sc *stab, const sc *uj, uint32_t lg, uint32_t n) {
    const uint32_t t = threadIdx.x;
    if (t == 0) {
        sc acc = sc_mont_one();
        for (uint32_t j = 0; j < lg; j++) acc = mm(acc, uj[lg + j]);
        stab[0] = acc;
    }
    __syncthreads();
    for (uint32_t b = 0; b < lg; b++) {
        const uint32_t half = 1u << b;
        sc usq = mm(uj[lg - 1 - b], uj[lg - 1 - b]);
        for (uint32_t i = half + t; i < 2 * half && i < n; i += BBP_SC_THREADS) stab[i] = mm(stab[i - half], usq);
        __syncthreads();
    }
}

// The same vector in product form: s[i] = s_lo[i mod 2^lo_bits] * s_hi[i >> lo_bits] (s is a product over the bits of i), so
// the verifier never materialises it: a * s[i] = s_lo[.] * (a s_hi)[.] is one product against the two (table build + use)
// of the recurrence. Each factor table is built by the same doubling recurrence on its own bits, in shared memory.
// s_lo: 2^lo_bits entries, s_hi: 2^(lg - lo_bits) entries. Synchronises the block.
__device__ inline void build_s_factors(sc *s_lo, sc *s_hi, const sc *uj, uint32_t lg, uint32_t lo_bits) {
    const uint32_t t = threadIdx.x, hi_bits = lg - lo_bits;
    if (t < 2) {
        // t = 0: bits 0 .. lo_bits-1 (challenges u_{lg-1} .. u_{lg-lo_bits}); t = 1: the rest
        sc acc = sc_mont_one();
        const uint32_t b0 = t ? lo_bits : 0, b1 = t ? lg : lo_bits;
        for (uint32_t b = b0; b < b1; b++) acc = mm(acc, uj[lg + (lg - 1 - b)]);
        (t ? s_hi : s_lo)[0] = acc;
    }
    __syncthreads();
    for (uint32_t b = 0; b < max(lo_bits, hi_bits); b++) {
        const uint32_t half = 1u << b;
        if (b < lo_bits) {
            sc usq = mm(uj[lg - 1 - b], uj[lg - 1 - b]);
            for (uint32_t i = half + t; i < 2 * half; i += BBP_SC_THREADS) s_lo[i] = mm(s_lo[i - half], usq);
        }
        if (b < hi_bits) {
            sc usq = mm(uj[lg - 1 - lo_bits - b], uj[lg - 1 - lo_bits - b]);
            for (uint32_t i = half + t; i < 2 * half; i += BBP_SC_THREADS) s_hi[i] = mm(s_hi[i - half], usq);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- verifier scalar assembly
// Per proof (weight rho, 1 for a single verification):
//   stat[0]   (B)          rho * ( w (t_x - a b) + r (x^2 (wc + delta) - t_x) )
//   stat[1]   (B_blinding) rho * ( -e_blinding - r t_x_blinding )
//   stat[2+i] (G[i])       rho * uf[i] (x y^-i wR[i] - a s[i])
//   stat[2+n+i] (H[i])     rho * uf[i] (y^-i (x wL[i] + wO[i] - b s[n-1-i]) - 1)
//   dyn_out[i] = rho * wV[i] r x^2   (coefficients of the V commitments)
// with s[i] = prod_j u_j^(+-1) (bit (lg n - 1 - j) of i set -> u_j, else u_j^-1), uf = 1 (i < n1) | u, delta = <y^-n wR, wL>.
// Results stay in Montgomery form; k_stat_reduce sums them over the batch and converts.
#define BBP_S_LO_BITS 8      // s_lo has min(n, 256) entries: with i = t + 256 k its index is constant per thread
#define BBP_S_MAX_HI 256     // n <= 65536
__global__ void __launch_bounds__(BBP_SC_THREADS, 3) k_verify_scalars(sc_batch B) {
    __shared__ sc smem[BBP_SC_THREADS];
    __shared__ sc uj[64];
    __shared__ sc long_val[BBP_MAX_LONG];
    __shared__ sc s_lo[1 << BBP_S_LO_BITS];
    extern __shared__ sc s_hi_all[];            // [5][n_hi]: s_hi, then rho a s_hi, rho b s_hi, rho u a s_hi, rho u b s_hi
    const uint32_t p = blockIdx.x, t = threadIdx.x, n = B.n, n1 = B.n1, lg = B.lg_n;
    const uint32_t lo_bits = min(lg, (uint32_t)BBP_S_LO_BITS), n_lo = 1u << lo_bits, n_hi = n >> lo_bits;
    const sc *ch = B.chal + (size_t)p * CH_N;
    const sc *zpow = B.zpow + (size_t)p * B.q, *yinv = B.yinvpow + (size_t)p * n;
    if (t < 2 * lg) uj[t] = sc_to_mont(ch[CH_UJ0 + t]);   // u_0..u_{lg-1}, then u_0^-1..u_{lg-1}^-1
    __syncthreads();
    const sc xM = sc_to_mont(ch[CH_X]), rhoM = sc_to_mont(ch[CH_RHO]);
    const uint32_t gc = B.gcols;
    sc *stat = B.stat + (size_t)p * (2 + 2 * gc);
    for (uint32_t i = n + t; i < gc; i += BBP_SC_THREADS) { stat[2 + i] = sc_zero(); stat[2 + gc + i] = sc_zero(); }
    // wc = - sum sign * z^(j+1) * pub[idx]
    sc wc = sc_zero();
    const sc *pub = B.pub + (size_t)p * B.n_pub;
    for (uint32_t e = t; e < B.n_const; e += BBP_SC_THREADS) {
        uint32_t v = B.const_j[e];
        sc term = mm(zpow[v & 0x7fffffffu], sc_to_mont(pub_value(B, pub, B.const_idx[e])));
        wc = (v >> 31) ? sc_add(wc, term) : sc_sub(wc, term);
    }
    wc = block_sum_sc(wc, smem);
    sc delta = sc_zero();
    flatten_long_rows(B, zpow, long_val, smem, pub);
    sc *s_hi = s_hi_all;
    build_s_factors(s_lo, s_hi, uj, lg, lo_bits);
    // rho, a / b and uf are folded into four copies of the (small) high factor table, so that a column costs 8 products
    // below n1 and 3 above:  g = (rho x) ywR - s_lo (rho a s_hi),  h = y^-i ((rho x) wL + rho wO - s_lo' (rho b s_hi')) - rho
    {
        const sc uM = sc_to_mont(ch[CH_U]), aM = sc_to_mont(ch[CH_A]), bM = sc_to_mont(ch[CH_B]);
        const sc ra = mm(rhoM, aM), rb = mm(rhoM, bM);
        for (uint32_t w = t; w < 4 * n_hi; w += BBP_SC_THREADS) {
            const uint32_t k = w / n_hi, h = w % n_hi;
            sc f = (k & 1) ? rb : ra;
            if (k >= 2) f = mm(f, uM);
            s_hi_all[(1 + k) * n_hi + h] = mm(f, s_hi[h]);
        }
    }
    __syncthreads();
    const sc rx = mm(rhoM, xM), ruU = mm(rhoM, sc_to_mont(ch[CH_U]));
    const sc *hi_a = s_hi_all + n_hi, *hi_b = s_hi_all + 2 * n_hi, *hi_au = s_hi_all + 3 * n_hi, *hi_bu = s_hi_all + 4 * n_hi;
    for (uint32_t i = t; i < n; i += BBP_SC_THREADS) {
        const uint32_t lo = i & (n_lo - 1), hi = i >> lo_bits;   // reversed index n - 1 - i: (n_lo - 1 - lo, n_hi - 1 - hi)
        sc g, h;
        if (i < n1) {
            sc wL = flatten_row_l(B, zpow, i, long_val, pub), wR = flatten_row_l(B, zpow, n1 + i, long_val, pub), wO = flatten_row_l(B, zpow, 2 * n1 + i, long_val, pub);
            const sc yi = yinv[i];
            sc ywR = mm(yi, wR);
            delta = sc_add(delta, mm(ywR, wL));
            g = sc_sub(mm(rx, ywR), mm(s_lo[lo], hi_a[hi]));
            h = sc_sub(mm(yi, sc_sub(sc_add(mm(rx, wL), mm(rhoM, wO)), mm(s_lo[n_lo - 1 - lo], hi_b[n_hi - 1 - hi]))), rhoM);
        } else {
            g = sc_neg(mm(s_lo[lo], hi_au[hi]));
            h = sc_sub(sc_neg(mm(yinv[i], mm(s_lo[n_lo - 1 - lo], hi_bu[n_hi - 1 - hi]))), ruU);
        }
        stat[2 + i] = g;
        stat[2 + gc + i] = h;
    }
    delta = block_sum_sc(delta, smem);
    sc rM = sc_to_mont(ch[CH_R]);
    sc x2 = mm(xM, xM), rx2 = mm(rM, x2);
    if (!B.dyn_done) {
        for (uint32_t i = t; i < B.m; i += BBP_SC_THREADS) {
            sc wV = flatten_row(B, zpow, 3 * n1 + i, pub);
            B.dyn_out[(size_t)p * B.dyn_stride + i] = sc_from_mont(mm(rhoM, mm(wV, rx2)));
        }
        // the transcript-dependent dynamic scalars (written unweighted by k_verify_transcript) take the batch weight here
        for (uint32_t i = B.m + t; i < B.dyn_stride; i += BBP_SC_THREADS) {
            sc *d = B.dyn_out + (size_t)p * B.dyn_stride + i;
            *d = mm(rhoM, *d);   // (rho R) * d / R = rho * d
        }
    }
    if (t == 0) {
        sc txM = sc_to_mont(ch[CH_TX]), wM = sc_to_mont(ch[CH_W]), aM = sc_to_mont(ch[CH_A]), bM = sc_to_mont(ch[CH_B]);
        sc bs = sc_add(mm(wM, sc_sub(txM, mm(aM, bM))), mm(rM, sc_sub(mm(x2, sc_add(wc, delta)), txM)));
        sc bbs = sc_sub(sc_neg(sc_to_mont(ch[CH_EBL])), mm(rM, sc_to_mont(ch[CH_TXBL])));
        stat[0] = mm(rhoM, bs);
        stat[1] = mm(rhoM, bbs);
    }
}
inline size_t k_verify_scalars_smem(uint32_t n, uint32_t lg) { return (size_t)5 * (n >> (lg < BBP_S_LO_BITS ? lg : BBP_S_LO_BITS)) * sizeof(sc); }

// The dynamic-point scalars alone, without the power tables: dyn_out[i] = rho wV[i] r x^2 for the m commitments (the few
// z powers a V row needs are assembled from z^(2^k) by the bits of the exponent) and rho * d for the transcript-dependent
// ones. Lets the variable-base MSM over a request's own points start on a second stream while k_powers / k_verify_scalars
// (which only feed the static-base MSM) are still running. One block per request.
__global__ void __launch_bounds__(64) k_dyn_weights(sc_batch B) {
    __shared__ sc z2k[16];
    const uint32_t p = blockIdx.x, t = threadIdx.x;
    const sc *ch = B.chal + (size_t)p * CH_N;
    if (t == 0) {
        sc r = sc_to_mont(ch[CH_Z]);
        for (uint32_t k = 0; k < 16; k++) { z2k[k] = r; r = mm(r, r); }
    }
    __syncthreads();
    const sc rhoM = sc_to_mont(ch[CH_RHO]);
    for (uint32_t i = t; i < B.dyn_stride; i += 64) {
        sc *d = B.dyn_out + (size_t)p * B.dyn_stride + i;
        if (i < B.m) {
            const uint32_t row = 3 * B.n1 + i;
            sc acc = sc_zero();
            for (uint32_t e = B.row_ptr[row]; e < B.row_ptr[row + 1]; e++) {
                const uint32_t v = B.entries[e], ex = (v & 0x7fffffffu) + 1;   // zpow[j] = z^(j+1)
                sc zp = sc_mont_one();
                for (uint32_t k = 0; k < 16; k++)
                    if ((ex >> k) & 1) zp = mm(zp, z2k[k]);
                if (B.coef) {
                    const uint32_t ci = B.coef[e];
                    if (ci) zp = mm(zp, sc_to_mont(pub_value(B, B.pub + (size_t)p * B.n_pub, ci)));
                }
                acc = (v >> 31) ? sc_sub(acc, zp) : sc_add(acc, zp);
            }
            const sc xM = sc_to_mont(ch[CH_X]);
            const sc rx2 = mm(sc_to_mont(ch[CH_R]), mm(xM, xM));
            *d = sc_from_mont(mm(rhoM, mm(acc, rx2)));
        } else {
            *d = mm(rhoM, *d);   // (rho R) * d / R = rho * d
        }
    }
}

// ================================================================ aggregated range proofs
// vectors are indexed k = j * bits + i (party j, bit i), the order bulletproofs' aggregated generators iterate in.
__device__ __forceinline__ uint32_t rp_bit(const sc_batch &B, uint32_t p, uint32_t k) {
    return (uint32_t)((B.rp_values[(size_t)p * B.rp_m + k / B.rp_bits] >> (k % B.rp_bits)) & 1ull);
}

// slot 2p: A = (sum a_blinding) B_bl + sum_k (bit ? G_k : -H_k);  slot 2p+1: S = (sum s_blinding) B_bl + <s_L, G> + <s_R, H>
__global__ void __launch_bounds__(BBP_SC_THREADS) k_rp_commit_slots(sc_batch B) {
    const uint32_t p = blockIdx.x / 2, which = blockIdx.x % 2, gc = B.gcols, slot_len = 2 + 2 * gc;
    sc *out = B.slots + (size_t)blockIdx.x * slot_len;
    sc minus_one = sc_neg(sc_one());
    for (uint32_t i = threadIdx.x; i < slot_len; i += BBP_SC_THREADS) {
        sc v = sc_zero();
        if (i == 1) v = B.blind3[(size_t)p * 3 + which];
        else if (i >= 2) {
            uint32_t k = (i - 2) % gc;
            bool is_h = (i - 2) >= gc;
            if (k < B.n) {
                if (which == 0) {
                    uint32_t bit = rp_bit(B, p, k);
                    if (!is_h && bit) v = sc_one();
                    if (is_h && !bit) v = minus_one;
                } else {
                    v = (is_h ? B.sR : B.sL)[(size_t)p * B.n + k];
                }
            }
        }
        out[i] = v;
    }
}

// l0 = a_L - z, l1 = s_L, r0 = y^k (a_R + z) + z^(2+j) 2^i, r1 = y^k s_R; t0 = <l0,r0>, t1 = <l0,r1> + <l1,r0>, t2 = <l1,r1>
__global__ void __launch_bounds__(BBP_SC_THREADS) k_rp_polys(sc_batch B) {
    __shared__ sc smem[BBP_SC_THREADS];
    const uint32_t p = blockIdx.x, t = threadIdx.x, n = B.n;
    const sc *ch = B.chal + (size_t)p * CH_N, *ypow = B.ypow + (size_t)p * n;
    sc *poly = B.poly + (size_t)p * 4 * n;
    sc zM = sc_to_mont(ch[CH_Z]), one = sc_mont_one();
    __shared__ sc zz[64], two[64];   // z^(2+j) for the first 64 parties, 2^i for the (at most 64) bits
    if (t < 64) { zz[t] = sc_pow_small_mont(zM, 2 + t); two[t] = sc_to_mont(sc_from_u64(1ull << t)); }
    __syncthreads();
    sc t0 = sc_zero(), t1 = t0, t2 = t0;
    for (uint32_t k = t; k < n; k += BBP_SC_THREADS) {
        uint32_t j = k / B.rp_bits, i = k % B.rp_bits, bit = rp_bit(B, p, k);
        sc aL = bit ? one : sc_zero(), aR = bit ? sc_zero() : sc_neg(one);
        sc sL = sc_to_mont(B.sL[(size_t)p * n + k]), sR = sc_to_mont(B.sR[(size_t)p * n + k]);
        sc zz_j = j < 64 ? zz[j] : sc_pow_small_mont(zM, 2 + j);
        sc two_i = two[i];
        sc l0 = sc_sub(aL, zM);
        sc r0 = sc_add(mm(ypow[k], sc_add(aR, zM)), mm(zz_j, two_i));
        sc r1 = mm(ypow[k], sR);
        poly[k] = l0; poly[n + k] = sL; poly[2 * n + k] = r0; poly[3 * n + k] = r1;
        t0 = sc_add(t0, mm(l0, r0));
        t1 = sc_add(t1, sc_add(mm(l0, r1), mm(sL, r0)));
        t2 = sc_add(t2, mm(sL, r1));
    }
    sc r[3] = {t0, t1, t2};
    for (int k = 0; k < 3; k++) {
        sc s = block_sum_sc(r[k], smem);
        if (t == 0) B.tout[(size_t)p * 8 + k] = sc_from_mont(s);
    }
}

// a = l0 + l1 x, b = r0 + r1 x, G factors 1, H factors y^-k
__global__ void __launch_bounds__(BBP_SC_THREADS) k_rp_ipp_init(sc_batch B) {
    const uint32_t p = blockIdx.x, n = B.n;
    const sc *ch = B.chal + (size_t)p * CH_N, *poly = B.poly + (size_t)p * 4 * n, *yinv = B.yinvpow + (size_t)p * n;
    sc xM = sc_to_mont(ch[CH_X]), one = sc_mont_one();
    for (uint32_t k = threadIdx.x; k < n; k += BBP_SC_THREADS) {
        size_t o = (size_t)p * n + k;
        B.a[o] = sc_add(poly[k], mm(poly[n + k], xM));
        B.b[o] = sc_add(poly[2 * n + k], mm(poly[3 * n + k], xM));
        B.sG[o] = one;
        B.sH[o] = yinv[k];
    }
}

// verifier: G_k gets rho (-z - a s[k]), H_k gets rho (z + y^-k (z^2 z^j 2^i - b s[n-1-k])); the B / B_blinding coefficients
// are computed on the host (they need only O(m + n) scalar work) and arrive, already weighted, in CH_TX / CH_TXBL.
__global__ void __launch_bounds__(BBP_SC_THREADS) k_rp_verify_scalars(sc_batch B) {
    __shared__ sc uj[64];
    const uint32_t p = blockIdx.x, t = threadIdx.x, n = B.n, lg = B.lg_n, gc = B.gcols;
    const sc *ch = B.chal + (size_t)p * CH_N, *yinv = B.yinvpow + (size_t)p * n;
    if (t < 2 * lg) uj[t] = sc_to_mont(ch[CH_UJ0 + t]);
    __syncthreads();
    sc zM = sc_to_mont(ch[CH_Z]), aM = sc_to_mont(ch[CH_A]), bM = sc_to_mont(ch[CH_B]), rhoM = sc_to_mont(ch[CH_RHO]), one = sc_mont_one();
    sc *stat = B.stat + (size_t)p * (2 + 2 * gc);
    for (uint32_t i = n + t; i < gc; i += BBP_SC_THREADS) { stat[2 + i] = sc_zero(); stat[2 + gc + i] = sc_zero(); }
    sc *stab = B.stab + (size_t)p * n;
    __shared__ sc zz[64], two[64];   // z^(2+j) for the first 64 parties, 2^i for the (at most 64) bits
    if (t < 64) { zz[t] = sc_pow_small_mont(zM, 2 + t); two[t] = sc_to_mont(sc_from_u64(1ull << t)); }
    build_s_table(stab, uj, lg, n);   // synchronises the block
    for (uint32_t k = t; k < n; k += BBP_SC_THREADS) {
        sc s = stab[k], srev = stab[n - 1 - k];
        uint32_t j = k / B.rp_bits, i = k % B.rp_bits;
        sc zz_j = j < 64 ? zz[j] : sc_pow_small_mont(zM, 2 + j);
        sc two_i = two[i];
        sc g = sc_sub(sc_neg(zM), mm(aM, s));
        sc h = sc_add(zM, mm(yinv[k], sc_sub(mm(zz_j, two_i), mm(bM, srev))));
        stat[2 + k] = mm(rhoM, g);
        stat[2 + gc + k] = mm(rhoM, h);
    }
    if (t == 0) { stat[0] = sc_to_mont(ch[CH_TX]); stat[1] = sc_to_mont(ch[CH_TXBL]); }
}

// sums the per-proof static-base scalars over groups of `group_size` consecutive proofs -> one slot of slot_len scalars
// per group (normal form). group_size = 1 converts each proof's slot for an individual check; group_size = n_proofs
// builds the single combined slot of a batch verification.
// total: rows in stat; the last group may be shorter than group_size
__global__ void __launch_bounds__(BBP_SC_THREADS) k_stat_reduce(const sc *__restrict__ stat, uint32_t group_size, uint32_t slot_len, sc *__restrict__ out,
                                                                uint32_t total = 0xffffffffu) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (i >= slot_len) return;
    sc acc = sc_zero();
    const sc *base = stat + (size_t)g * group_size * slot_len;
    const uint32_t cnt = min(group_size, total - min(total, g * group_size));
    for (uint32_t p = 0; p < cnt; p++) acc = sc_add(acc, base[(size_t)p * slot_len + i]);
    out[(size_t)g * slot_len + i] = sc_from_mont(acc);
}

// the same for large groups: 16 columns x 16 proof slices per block, so that a batch-wide sum is not 4098 threads walking
// the whole batch one proof at a time
__global__ void __launch_bounds__(256) k_stat_reduce_wide(const sc *__restrict__ stat, uint32_t group_size, uint32_t slot_len, sc *__restrict__ out,
                                                          uint32_t total = 0xffffffffu) {
    __shared__ sc part[16][16];
    const uint32_t col = threadIdx.x & 15, slice = threadIdx.x >> 4, i = blockIdx.x * 16 + col, g = blockIdx.y;
    sc acc = sc_zero();
    if (i < slot_len) {
        const sc *base = stat + (size_t)g * group_size * slot_len + i;
        const uint32_t cnt = min(group_size, total - min(total, g * group_size));
        for (uint32_t p = slice; p < cnt; p += 16) acc = sc_add(acc, base[(size_t)p * slot_len]);
    }
    part[slice][col] = acc;
    __syncthreads();
    for (uint32_t s = 8; s >= 1; s >>= 1) {
        if (slice < s) part[slice][col] = sc_add(part[slice][col], part[slice + s][col]);
        __syncthreads();
    }
    if (slice == 0 && i < slot_len) out[(size_t)g * slot_len + i] = sc_from_mont(part[0][col]);
}

}  // namespace bbp
