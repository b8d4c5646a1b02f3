// ORACLE — test infrastructure only (see fe.h header).
// bulletproofs 1.0.4 aggregated range proofs: RangeProof::{prove_multiple, verify_multiple} with the
// dealer/party MPC run in-process. The reference itself has no call site;
// BASELINE.json configs[4] names it (m = 64 parties x n = 64 bits => 4096-point IPP).
// PARITY STATUS: proof-byte level parity UNPINNED (no reference vectors; see r1cs.h). 
// RNG contract: upstream draws from the caller's rng directly (not a TranscriptRng): every party owns an rng
// (Party::new(.., rng)). Here party j's rng is the SHAKE256 stream of (32-byte seed || v_blinding_j (32 B canonical) ||
// LE64(value_j) || LE32(j)) — keyed with the party's witness, so that a seed reused for a different value set does not
// repeat the nonces — consumed 64 bytes per Scalar::random in upstream's per-party draw order: a_blinding, s_blinding,
// s_L[0..n), s_R[0..n), then (after the y, z challenges) t_1_blinding, t_2_blinding. Independent per-party streams are
// what lets the product squeeze them in parallel. The verifier's batching scalar c is the first scalar of
// transcript.build_rng().finalize(rng32) taken AFTER the whole proof (all L_j, R_j included) has been absorbed, so it is
// bound to every proof byte as well as to the caller's secret seed.
#pragma once
#include "gens.h"
#include "ipp.h"
#include "merlin.h"
#include "msm.h"

namespace orc {

struct shake_rng {
    shake256 s;
    explicit shake_rng(const uint8_t seed[32]) { s.absorb(seed, 32); }
    shake_rng(const uint8_t seed[32], const sc &v_blinding, uint64_t value, uint32_t party) {
        uint8_t bl[32], tail[12];
        sc_tobytes(bl, v_blinding);
        for (int i = 0; i < 8; i++) tail[i] = (uint8_t)(value >> (8 * i));
        for (int i = 0; i < 4; i++) tail[8 + i] = (uint8_t)(party >> (8 * i));
        s.absorb(seed, 32);
        s.absorb(bl, 32);
        s.absorb(tail, 12);
    }
    sc random_scalar() {
        uint8_t b[64];
        s.squeeze(b, 64);
        return sc_from_wide(b);
    }
};

static inline sc sc_pow_u64(const sc &x, uint64_t e) {
    sc r = sc_one(), b = x;
    while (e) {
        if (e & 1) r = sc_mul(r, b);
        b = sc_mul(b, b);
        e >>= 1;
    }
    return r;
}
static inline sc sc_sum_of_powers(const sc &x, size_t n) {  // 1 + x + ... + x^(n-1)
    sc acc = sc_zero(), e = sc_one();
    for (size_t i = 0; i < n; i++) { acc = sc_add(acc, e); e = sc_mul(e, x); }
    return acc;
}

static inline int rangeproof_prove_multiple(const std::vector<uint64_t> &values, const std::vector<sc> &blindings, size_t n,
                                            const uint8_t rng_seed[32], std::vector<uint8_t> &proof_out, std::vector<bytes32> &V_out) {
    size_t m = values.size();
    if (!(n == 8 || n == 16 || n == 32 || n == 64)) return -1;
    if (m == 0 || (m & (m - 1)) != 0 || blindings.size() != m) return -1;
    pedersen_gens pc;
    bulletproof_gens bp(n, m);
    transcript tr("bbp-rangeproof");   // caller-chosen label; the product uses the same one
    std::vector<shake_rng> rngs;
    for (size_t j = 0; j < m; j++) rngs.emplace_back(rng_seed, blindings[j], values[j], (uint32_t)j);

    tr.rangeproof_domain_sep(n, m);
    struct party { sc a_bl, s_bl; std::vector<sc> s_L, s_R; ge A, S; sc t1_bl, t2_bl, t0, t1, t2; std::vector<sc> l0, l1, r0, r1; };
    std::vector<party> P(m);
    V_out.resize(m);
    for (size_t j = 0; j < m; j++) {
        party &p = P[j];
        V_out[j] = ge_compress32(pc.commit(sc_from_u64(values[j]), blindings[j]));
        p.a_bl = rngs[j].random_scalar();
        ge A = ge_scalarmul(p.a_bl, pc.B_blinding);
        for (size_t i = 0; i < n; i++) {
            if ((values[j] >> i) & 1) A = ge_add(A, bp.G[j][i]);
            else A = ge_sub(A, bp.H[j][i]);
        }
        p.A = A;
        p.s_bl = rngs[j].random_scalar();
        p.s_L.resize(n); p.s_R.resize(n);
        for (size_t i = 0; i < n; i++) p.s_L[i] = rngs[j].random_scalar();
        for (size_t i = 0; i < n; i++) p.s_R[i] = rngs[j].random_scalar();
        std::vector<sc> ss; std::vector<ge> pp;
        ss.push_back(p.s_bl); pp.push_back(pc.B_blinding);
        for (size_t i = 0; i < n; i++) { ss.push_back(p.s_L[i]); pp.push_back(bp.G[j][i]); }
        for (size_t i = 0; i < n; i++) { ss.push_back(p.s_R[i]); pp.push_back(bp.H[j][i]); }
        p.S = msm_pippenger_serial(ss.data(), pp.data(), ss.size());
    }
    for (size_t j = 0; j < m; j++) tr.append_point("V", V_out[j].data());
    ge A = ge_identity(), S = ge_identity();
    for (size_t j = 0; j < m; j++) { A = ge_add(A, P[j].A); S = ge_add(S, P[j].S); }
    bytes32 Ac = ge_compress32(A), Sc = ge_compress32(S);
    tr.append_point("A", Ac.data());
    tr.append_point("S", Sc.data());
    sc y = tr.challenge_scalar("y"), z = tr.challenge_scalar("z");

    ge T1 = ge_identity(), T2 = ge_identity();
    sc zz = sc_mul(z, z);
    for (size_t j = 0; j < m; j++) {
        party &p = P[j];
        sc offset_y = sc_pow_u64(y, (uint64_t)(j * n)), offset_z = sc_pow_u64(z, (uint64_t)j);
        sc offset_zz = sc_mul(zz, offset_z);
        p.l0.resize(n); p.l1.resize(n); p.r0.resize(n); p.r1.resize(n);
        sc exp_y = offset_y, exp_2 = sc_one();
        for (size_t i = 0; i < n; i++) {
            sc a_L = sc_from_u64((values[j] >> i) & 1);
            sc a_R = sc_sub(a_L, sc_one());
            p.l0[i] = sc_sub(a_L, z);
            p.l1[i] = p.s_L[i];
            p.r0[i] = sc_add(sc_mul(exp_y, sc_add(a_R, z)), sc_mul(offset_zz, exp_2));
            p.r1[i] = sc_mul(exp_y, p.s_R[i]);
            exp_y = sc_mul(exp_y, y);
            exp_2 = sc_add(exp_2, exp_2);
        }
        p.t0 = sc_inner_product(p.l0.data(), p.r0.data(), n);
        p.t2 = sc_inner_product(p.l1.data(), p.r1.data(), n);
        p.t1 = sc_add(sc_inner_product(p.l0.data(), p.r1.data(), n), sc_inner_product(p.l1.data(), p.r0.data(), n));
        p.t1_bl = rngs[j].random_scalar();
        p.t2_bl = rngs[j].random_scalar();
        T1 = ge_add(T1, pc.commit(p.t1, p.t1_bl));
        T2 = ge_add(T2, pc.commit(p.t2, p.t2_bl));
    }
    bytes32 T1c = ge_compress32(T1), T2c = ge_compress32(T2);
    tr.append_point("T_1", T1c.data());
    tr.append_point("T_2", T2c.data());
    sc x = tr.challenge_scalar("x");
    if (sc_iszero(x)) return -4;   // MaliciousDealer

    sc t_x = sc_zero(), t_x_bl = sc_zero(), e_bl = sc_zero();
    std::vector<sc> l_vec(n * m), r_vec(n * m);
    for (size_t j = 0; j < m; j++) {
        party &p = P[j];
        sc offset_zz = sc_mul(zz, sc_pow_u64(z, (uint64_t)j));
        t_x = sc_add(t_x, sc_add(p.t0, sc_mul(x, sc_add(p.t1, sc_mul(x, p.t2)))));
        t_x_bl = sc_add(t_x_bl, sc_add(sc_mul(offset_zz, blindings[j]), sc_mul(x, sc_add(p.t1_bl, sc_mul(x, p.t2_bl)))));
        e_bl = sc_add(e_bl, sc_add(p.a_bl, sc_mul(p.s_bl, x)));
        for (size_t i = 0; i < n; i++) {
            l_vec[j * n + i] = sc_add(p.l0[i], sc_mul(p.l1[i], x));
            r_vec[j * n + i] = sc_add(p.r0[i], sc_mul(p.r1[i], x));
        }
    }
    tr.append_scalar("t_x", t_x);
    tr.append_scalar("t_x_blinding", t_x_bl);
    tr.append_scalar("e_blinding", e_bl);
    sc w = tr.challenge_scalar("w");
    ge Q = ge_scalarmul(w, pc.B);
    std::vector<sc> Gf(n * m, sc_one()), Hf(n * m);
    sc y_inv = sc_invert(y), e = sc_one();
    for (size_t i = 0; i < n * m; i++) { Hf[i] = e; e = sc_mul(e, y_inv); }
    std::vector<ge> G, H;
    for (size_t j = 0; j < m; j++) {
        G.insert(G.end(), bp.G[j].begin(), bp.G[j].begin() + n);
        H.insert(H.end(), bp.H[j].begin(), bp.H[j].begin() + n);
    }
    ipp_proof ipp = ipp_create(tr, Q, Gf, Hf, std::move(G), std::move(H), std::move(l_vec), std::move(r_vec));

    proof_out.clear();
    auto put = [&](const bytes32 &b) { proof_out.insert(proof_out.end(), b.begin(), b.end()); };
    auto puts = [&](const sc &s) { uint8_t t[32]; sc_tobytes(t, s); proof_out.insert(proof_out.end(), t, t + 32); };
    put(Ac); put(Sc); put(T1c); put(T2c); puts(t_x); puts(t_x_bl); puts(e_bl);
    std::vector<uint8_t> ib = ipp_to_bytes(ipp);
    proof_out.insert(proof_out.end(), ib.begin(), ib.end());
    return 0;
}

// 0 = accept; -2 format error; -3 verification error; -1 bad parameters
static inline int rangeproof_verify_multiple(const uint8_t *proof, size_t len, const std::vector<bytes32> &V, size_t n, const uint8_t rng32[32],
                                             int threads = 1) {
    size_t m = V.size();
    if (!(n == 8 || n == 16 || n == 32 || n == 64)) return -1;
    if (m == 0 || (m & (m - 1)) != 0) return -1;
    if (len % 32 != 0 || len < 7 * 32) return -2;
    bytes32 A, S, T1, T2;
    memcpy(A.data(), proof, 32); memcpy(S.data(), proof + 32, 32); memcpy(T1.data(), proof + 64, 32); memcpy(T2.data(), proof + 96, 32);
    sc t_x, t_x_bl, e_bl;
    if (!sc_from_canonical(t_x, proof + 128) || !sc_from_canonical(t_x_bl, proof + 160) || !sc_from_canonical(e_bl, proof + 192)) return -2;
    ipp_proof ipp;
    if (!ipp_from_bytes(ipp, proof + 224, len - 224)) return -2;

    pedersen_gens pc;
    bulletproof_gens bp(n, m);
    transcript tr("bbp-rangeproof");
    tr.rangeproof_domain_sep(n, m);
    for (size_t j = 0; j < m; j++) tr.append_point("V", V[j].data());
    if (!tr.validate_and_append_point("A", A.data())) return -3;
    if (!tr.validate_and_append_point("S", S.data())) return -3;
    sc y = tr.challenge_scalar("y"), z = tr.challenge_scalar("z");
    sc zz = sc_mul(z, z), minus_z = sc_neg(z);
    if (!tr.validate_and_append_point("T_1", T1.data())) return -3;
    if (!tr.validate_and_append_point("T_2", T2.data())) return -3;
    sc x = tr.challenge_scalar("x");
    tr.append_scalar("t_x", t_x);
    tr.append_scalar("t_x_blinding", t_x_bl);
    tr.append_scalar("e_blinding", e_bl);
    sc w = tr.challenge_scalar("w");
    std::vector<sc> x_sq, x_inv_sq, s;
    if (!ipp_verification_scalars(ipp, n * m, tr, x_sq, x_inv_sq, s)) return -3;
    sc c;
    {
        transcript_rng r = tr.build_rng().finalize(rng32);
        c = r.random_scalar();
    }
    sc a = ipp.a, b = ipp.b;
    size_t nm = n * m;

    std::vector<sc> ms;
    std::vector<ge> mp;
    bool ok = true;
    auto push_c = [&](const sc &sv, const bytes32 &pt) {
        ms.push_back(sv);
        ge p;
        if (ok && !ge_decompress(p, pt.data())) ok = false;
        if (ok) mp.push_back(p);
    };
    push_c(sc_one(), A);
    push_c(x, S);
    push_c(sc_mul(c, x), T1);
    push_c(sc_mul(c, sc_mul(x, x)), T2);
    for (size_t i = 0; i < x_sq.size(); i++) push_c(x_sq[i], ipp.L[i]);
    for (size_t i = 0; i < x_inv_sq.size(); i++) push_c(x_inv_sq[i], ipp.R[i]);
    if (!ok) return -3;
    ms.push_back(sc_sub(sc_neg(e_bl), sc_mul(c, t_x_bl))); mp.push_back(pc.B_blinding);
    sc sum_y = sc_sum_of_powers(y, nm), sum_2 = sc_sum_of_powers(sc_from_u64(2), n), sum_z = sc_sum_of_powers(z, m);
    sc delta = sc_sub(sc_mul(sc_sub(z, zz), sum_y), sc_mul(sc_mul(sc_mul(zz, z), sum_2), sum_z));
    ms.push_back(sc_add(sc_mul(w, sc_sub(t_x, sc_mul(a, b))), sc_mul(c, sc_sub(delta, t_x)))); mp.push_back(pc.B);
    for (size_t i = 0; i < nm; i++) { ms.push_back(sc_sub(minus_z, sc_mul(a, s[i]))); mp.push_back(bp.G[i / n][i % n]); }
    {
        sc y_inv = sc_invert(y), exp_y_inv = sc_one();
        sc exp_z = sc_one();
        for (size_t j = 0; j < m; j++) {
            sc exp_2 = sc_one();
            for (size_t i = 0; i < n; i++) {
                size_t k = j * n + i;
                sc z_and_2 = sc_mul(exp_2, exp_z);
                ms.push_back(sc_add(z, sc_mul(exp_y_inv, sc_sub(sc_mul(zz, z_and_2), sc_mul(b, s[nm - 1 - k])))));
                mp.push_back(bp.H[j][i]);
                exp_y_inv = sc_mul(exp_y_inv, y_inv);
                exp_2 = sc_add(exp_2, exp_2);
            }
            exp_z = sc_mul(exp_z, z);
        }
    }
    {
        sc exp_z = sc_one();
        for (size_t j = 0; j < m; j++) {
            push_c(sc_mul(sc_mul(c, zz), exp_z), V[j]);
            exp_z = sc_mul(exp_z, z);
        }
    }
    if (!ok) return -3;
    ge mega = msm_pippenger(ms.data(), mp.data(), ms.size(), threads);
    return ge_is_identity(mega) ? 0 : -3;
}

}  // namespace orc
