// ORACLE — test infrastructure only (see fe.h header).
// bulletproofs@4a05305 (branch develop, feature yoloproofs; Cargo.toml:23-26, Cargo.lock:65-67) R1CS
// proof system: ConstraintSystem / LinearCombination / Variable / Prover / Verifier / R1CSProof
// (SURVEY.md §2.2 U6, §8 a-4, a-5, a-7). Call sites in the reference: src/blindbid/proof.rs:50-88,
// src/blindbid/verify.rs:51-88, src/gadgets.rs (multiply / constrain).
// Only the one-phase (no randomized constraints) flow the reference circuit exercises is restated.
// PARITY STATUS: the group / field / scalar / Merlin layers underneath are pinned (RFC 9496, libsodium, Merlin's vector:
// tests/test_oracle_primitives.py); at PROOF-BYTE level parity is UNPINNED — the reference ships no tests or vectors, its
// Rust dependencies cannot be built here, so transcript labels, draw order and the R1CSProof byte layout are restated from
// the pinned upstream revision and checked only for self-consistency (SURVEY.md §8c risks R1-R4).
// RNG contract: the 32 "external" bytes that the reference draws from thread_rng in
// TranscriptRngBuilder::finalize are an explicit argument.
#pragma once
#include "gens.h"
#include "ipp.h"
#include "merlin.h"
#include "msm.h"
#include <utility>
#include <vector>

namespace orc {

enum r1cs_error { R1CS_OK = 0, R1CS_INVALID_GENERATORS_LENGTH = -1, R1CS_FORMAT_ERROR = -2, R1CS_VERIFICATION_ERROR = -3 };

enum var_kind : uint8_t { VAR_COMMITTED = 0, VAR_MUL_LEFT = 1, VAR_MUL_RIGHT = 2, VAR_MUL_OUT = 3, VAR_ONE = 4 };
struct variable {
    var_kind kind;
    uint32_t idx;
};
static inline variable var_one() { return variable{VAR_ONE, 0}; }

struct lincomb {
    std::vector<std::pair<variable, sc>> terms;
    lincomb() {}
    lincomb(const variable &v) { terms.push_back({v, sc_one()}); }
    lincomb(const sc &s) { terms.push_back({var_one(), s}); }
};
static inline lincomb operator+(lincomb a, const lincomb &b) {
    a.terms.insert(a.terms.end(), b.terms.begin(), b.terms.end());
    return a;
}
static inline lincomb operator-(lincomb a, const lincomb &b) {
    for (auto &t : b.terms) a.terms.push_back({t.first, sc_neg(t.second)});
    return a;
}

struct constraint_system {
    virtual ~constraint_system() {}
    virtual void multiply(lincomb left, lincomb right, variable &l, variable &r, variable &o) = 0;
    virtual void constrain(lincomb lc) = 0;
};

struct r1cs_proof {
    bytes32 A_I1, A_O1, S1, A_I2, A_O2, S2, T_1, T_3, T_4, T_5, T_6;
    sc t_x, t_x_blinding, e_blinding;
    ipp_proof ipp;
};

static inline bool bytes32_is_zero(const bytes32 &b) {
    uint8_t acc = 0;
    for (auto x : b) acc |= x;
    return acc == 0;
}

// R1CSProof::to_bytes. versioned = true: develop-branch layout with the leading phase byte
// (0 = one-phase, A_I2/A_O2/S2 omitted); false: legacy 14-point layout (SURVEY.md §8c risk R1).
static inline std::vector<uint8_t> r1cs_proof_to_bytes(const r1cs_proof &p, bool versioned = true) {
    std::vector<uint8_t> out;
    auto put = [&](const bytes32 &b) { out.insert(out.end(), b.begin(), b.end()); };
    auto puts = [&](const sc &s) { uint8_t t[32]; sc_tobytes(t, s); out.insert(out.end(), t, t + 32); };
    bool one_phase = bytes32_is_zero(p.A_I2) && bytes32_is_zero(p.A_O2) && bytes32_is_zero(p.S2);
    if (versioned) out.push_back(one_phase ? 0 : 1);
    put(p.A_I1); put(p.A_O1); put(p.S1);
    if (!versioned || !one_phase) { put(p.A_I2); put(p.A_O2); put(p.S2); }
    put(p.T_1); put(p.T_3); put(p.T_4); put(p.T_5); put(p.T_6);
    puts(p.t_x); puts(p.t_x_blinding); puts(p.e_blinding);
    std::vector<uint8_t> ib = ipp_to_bytes(p.ipp);
    out.insert(out.end(), ib.begin(), ib.end());
    return out;
}

static inline int r1cs_proof_from_bytes(r1cs_proof &p, const uint8_t *in, size_t len, bool versioned = true) {
    int version = 1;
    if (versioned) {
        if (len < 1) return R1CS_FORMAT_ERROR;
        version = in[0];
        in++; len--;
    }
    if (len % 32 != 0) return R1CS_FORMAT_ERROR;
    size_t minlen;
    if (version == 0) minlen = 11 * 32;
    else if (version == 1) minlen = 14 * 32;
    else return R1CS_FORMAT_ERROR;
    if (len < minlen) return R1CS_FORMAT_ERROR;
    size_t pos = 0;
    auto get = [&](bytes32 &b) { memcpy(b.data(), in + pos, 32); pos += 32; };
    get(p.A_I1); get(p.A_O1); get(p.S1);
    if (version == 0) { p.A_I2.fill(0); p.A_O2.fill(0); p.S2.fill(0); }
    else { get(p.A_I2); get(p.A_O2); get(p.S2); }
    get(p.T_1); get(p.T_3); get(p.T_4); get(p.T_5); get(p.T_6);
    if (!sc_from_canonical(p.t_x, in + pos)) return R1CS_FORMAT_ERROR;
    pos += 32;
    if (!sc_from_canonical(p.t_x_blinding, in + pos)) return R1CS_FORMAT_ERROR;
    pos += 32;
    if (!sc_from_canonical(p.e_blinding, in + pos)) return R1CS_FORMAT_ERROR;
    pos += 32;
    if (!ipp_from_bytes(p.ipp, in + pos, len - pos)) return R1CS_FORMAT_ERROR;
    return R1CS_OK;
}

static inline size_t next_pow2(size_t n) {
    size_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

struct prover : constraint_system {
    const pedersen_gens &pc;
    transcript &tr;
    std::vector<lincomb> constraints;
    std::vector<sc> a_L, a_R, a_O, v, v_blinding;

    prover(const pedersen_gens &pc_gens, transcript &t) : pc(pc_gens), tr(t) { tr.r1cs_domain_sep(); }

    // Prover::commit (proof.rs:55-67 call sites)
    variable commit(const sc &val, const sc &blinding, bytes32 &V_out) {
        uint32_t i = (uint32_t)v.size();
        v.push_back(val);
        v_blinding.push_back(blinding);
        V_out = ge_compress32(pc.commit(val, blinding));
        tr.append_point("V", V_out.data());
        return variable{VAR_COMMITTED, i};
    }
    sc eval(const lincomb &lc) const {
        sc acc = sc_zero();
        for (auto &t : lc.terms) {
            sc val;
            switch (t.first.kind) {
                case VAR_COMMITTED: val = v[t.first.idx]; break;
                case VAR_MUL_LEFT: val = a_L[t.first.idx]; break;
                case VAR_MUL_RIGHT: val = a_R[t.first.idx]; break;
                case VAR_MUL_OUT: val = a_O[t.first.idx]; break;
                default: val = sc_one(); break;
            }
            acc = sc_add(acc, sc_mul(t.second, val));
        }
        return acc;
    }
    void multiply(lincomb left, lincomb right, variable &l, variable &r, variable &o) override {
        sc lv = eval(left), rv = eval(right), ov = sc_mul(lv, rv);
        l = variable{VAR_MUL_LEFT, (uint32_t)a_L.size()};
        r = variable{VAR_MUL_RIGHT, (uint32_t)a_R.size()};
        o = variable{VAR_MUL_OUT, (uint32_t)a_O.size()};
        a_L.push_back(lv); a_R.push_back(rv); a_O.push_back(ov);
        left.terms.push_back({l, sc_neg(sc_one())});
        right.terms.push_back({r, sc_neg(sc_one())});
        constrain(std::move(left));
        constrain(std::move(right));
    }
    void constrain(lincomb lc) override { constraints.push_back(std::move(lc)); }

    void flattened_constraints(const sc &z, std::vector<sc> &wL, std::vector<sc> &wR, std::vector<sc> &wO, std::vector<sc> &wV) const {
        size_t n = a_L.size(), m = v.size();
        wL.assign(n, sc_zero()); wR.assign(n, sc_zero()); wO.assign(n, sc_zero()); wV.assign(m, sc_zero());
        sc exp_z = z;
        for (auto &lc : constraints) {
            for (auto &t : lc.terms) {
                sc term = sc_mul(exp_z, t.second);
                switch (t.first.kind) {
                    case VAR_MUL_LEFT: wL[t.first.idx] = sc_add(wL[t.first.idx], term); break;
                    case VAR_MUL_RIGHT: wR[t.first.idx] = sc_add(wR[t.first.idx], term); break;
                    case VAR_MUL_OUT: wO[t.first.idx] = sc_add(wO[t.first.idx], term); break;
                    case VAR_COMMITTED: wV[t.first.idx] = sc_sub(wV[t.first.idx], term); break;
                    default: break;  // the prover ignores constant terms
                }
            }
            exp_z = sc_mul(exp_z, z);
        }
    }

    // Prover::prove (proof.rs:88). Returns an r1cs_error.
    int prove(const bulletproof_gens &bp, const uint8_t external_rng32[32], r1cs_proof &out) {
        tr.append_u64("m", v.size());
        transcript_rng_builder rb = tr.build_rng();
        for (auto &vb : v_blinding) {
            uint8_t b[32];
            sc_tobytes(b, vb);
            rb.rekey_with_witness_bytes("v_blinding", b, 32);
        }
        transcript_rng rng = rb.finalize(external_rng32);

        size_t n1 = a_L.size();
        if (bp.gens_capacity < n1) return R1CS_INVALID_GENERATORS_LENGTH;
        const std::vector<ge> &G = bp.G[0], &H = bp.H[0];

        sc i_bl1 = rng.random_scalar(), o_bl1 = rng.random_scalar(), s_bl1 = rng.random_scalar();
        std::vector<sc> s_L(n1), s_R(n1);
        for (size_t i = 0; i < n1; i++) s_L[i] = rng.random_scalar();
        for (size_t i = 0; i < n1; i++) s_R[i] = rng.random_scalar();

        {
            std::vector<sc> ss; std::vector<ge> pp;
            ss.push_back(i_bl1); pp.push_back(pc.B_blinding);
            for (size_t i = 0; i < n1; i++) { ss.push_back(a_L[i]); pp.push_back(G[i]); }
            for (size_t i = 0; i < n1; i++) { ss.push_back(a_R[i]); pp.push_back(H[i]); }
            out.A_I1 = ge_compress32(msm_pippenger_serial(ss.data(), pp.data(), ss.size()));
            ss.clear(); pp.clear();
            ss.push_back(o_bl1); pp.push_back(pc.B_blinding);
            for (size_t i = 0; i < n1; i++) { ss.push_back(a_O[i]); pp.push_back(G[i]); }
            out.A_O1 = ge_compress32(msm_pippenger_serial(ss.data(), pp.data(), ss.size()));
            ss.clear(); pp.clear();
            ss.push_back(s_bl1); pp.push_back(pc.B_blinding);
            for (size_t i = 0; i < n1; i++) { ss.push_back(s_L[i]); pp.push_back(G[i]); }
            for (size_t i = 0; i < n1; i++) { ss.push_back(s_R[i]); pp.push_back(H[i]); }
            out.S1 = ge_compress32(msm_pippenger_serial(ss.data(), pp.data(), ss.size()));
        }
        tr.append_point("A_I1", out.A_I1.data());
        tr.append_point("A_O1", out.A_O1.data());
        tr.append_point("S1", out.S1.data());

        // no randomized (second-phase) constraints in the reference circuit
        tr.r1cs_1phase_domain_sep();
        size_t n = a_L.size();
        size_t padded_n = next_pow2(n);
        size_t pad = padded_n - n;
        if (bp.gens_capacity < padded_n) return R1CS_INVALID_GENERATORS_LENGTH;
        out.A_I2.fill(0); out.A_O2.fill(0); out.S2.fill(0);
        tr.append_point("A_I2", out.A_I2.data());
        tr.append_point("A_O2", out.A_O2.data());
        tr.append_point("S2", out.S2.data());

        sc y = tr.challenge_scalar("y");
        sc z = tr.challenge_scalar("z");
        std::vector<sc> wL, wR, wO, wV;
        flattened_constraints(z, wL, wR, wO, wV);

        std::vector<sc> l1(n), l2(n), l3(n), r0(n), r1(n), r3(n);
        sc y_inv = sc_invert(y);
        std::vector<sc> exp_y_inv(padded_n);
        {
            sc e = sc_one();
            for (size_t i = 0; i < padded_n; i++) { exp_y_inv[i] = e; e = sc_mul(e, y_inv); }
        }
        sc exp_y = sc_one();
        for (size_t i = 0; i < n; i++) {
            l1[i] = sc_add(a_L[i], sc_mul(exp_y_inv[i], wR[i]));
            l2[i] = a_O[i];
            l3[i] = s_L[i];
            r0[i] = sc_sub(wO[i], exp_y);
            r1[i] = sc_add(sc_mul(exp_y, a_R[i]), wL[i]);
            r3[i] = sc_mul(exp_y, s_R[i]);
            exp_y = sc_mul(exp_y, y);
        }
        sc t1 = sc_inner_product(l1.data(), r0.data(), n);
        sc t2 = sc_add(sc_inner_product(l1.data(), r1.data(), n), sc_inner_product(l2.data(), r0.data(), n));
        sc t3 = sc_add(sc_inner_product(l2.data(), r1.data(), n), sc_inner_product(l3.data(), r0.data(), n));
        sc t4 = sc_add(sc_inner_product(l1.data(), r3.data(), n), sc_inner_product(l3.data(), r1.data(), n));
        sc t5 = sc_inner_product(l2.data(), r3.data(), n);
        sc t6 = sc_inner_product(l3.data(), r3.data(), n);

        sc t1_bl = rng.random_scalar(), t3_bl = rng.random_scalar(), t4_bl = rng.random_scalar(),
           t5_bl = rng.random_scalar(), t6_bl = rng.random_scalar();
        out.T_1 = ge_compress32(pc.commit(t1, t1_bl));
        out.T_3 = ge_compress32(pc.commit(t3, t3_bl));
        out.T_4 = ge_compress32(pc.commit(t4, t4_bl));
        out.T_5 = ge_compress32(pc.commit(t5, t5_bl));
        out.T_6 = ge_compress32(pc.commit(t6, t6_bl));
        tr.append_point("T_1", out.T_1.data());
        tr.append_point("T_3", out.T_3.data());
        tr.append_point("T_4", out.T_4.data());
        tr.append_point("T_5", out.T_5.data());
        tr.append_point("T_6", out.T_6.data());

        sc u = tr.challenge_scalar("u");
        sc x = tr.challenge_scalar("x");

        sc t2_bl = sc_zero();
        for (size_t i = 0; i < wV.size(); i++) t2_bl = sc_add(t2_bl, sc_mul(wV[i], v_blinding[i]));

        auto poly6 = [&](const sc &c1, const sc &c2, const sc &c3, const sc &c4, const sc &c5, const sc &c6) {
            sc acc = c6;
            acc = sc_add(c5, sc_mul(x, acc));
            acc = sc_add(c4, sc_mul(x, acc));
            acc = sc_add(c3, sc_mul(x, acc));
            acc = sc_add(c2, sc_mul(x, acc));
            acc = sc_add(c1, sc_mul(x, acc));
            return sc_mul(x, acc);
        };
        out.t_x = poly6(t1, t2, t3, t4, t5, t6);
        out.t_x_blinding = poly6(t1_bl, t2_bl, t3_bl, t4_bl, t5_bl, t6_bl);

        std::vector<sc> l_vec(padded_n, sc_zero()), r_vec(padded_n, sc_zero());
        sc x2 = sc_mul(x, x), x3 = sc_mul(x2, x);
        for (size_t i = 0; i < n; i++) {
            l_vec[i] = sc_add(sc_add(sc_mul(l1[i], x), sc_mul(l2[i], x2)), sc_mul(l3[i], x3));
            r_vec[i] = sc_add(sc_add(r0[i], sc_mul(r1[i], x)), sc_mul(r3[i], x3));
        }
        for (size_t i = n; i < padded_n; i++) {
            r_vec[i] = sc_neg(exp_y);
            exp_y = sc_mul(exp_y, y);
        }
        // i_blinding2 = o_blinding2 = s_blinding2 = 0 in the one-phase case
        out.e_blinding = sc_mul(x, sc_add(i_bl1, sc_mul(x, sc_add(o_bl1, sc_mul(x, s_bl1)))));

        tr.append_scalar("t_x", out.t_x);
        tr.append_scalar("t_x_blinding", out.t_x_blinding);
        tr.append_scalar("e_blinding", out.e_blinding);

        sc w = tr.challenge_scalar("w");
        ge Q = ge_scalarmul(w, pc.B);

        std::vector<sc> Gf(padded_n), Hf(padded_n);
        for (size_t i = 0; i < padded_n; i++) {
            Gf[i] = (i < n1) ? sc_one() : u;
            Hf[i] = sc_mul(exp_y_inv[i], Gf[i]);
        }
        std::vector<ge> Gv(G.begin(), G.begin() + padded_n), Hv(H.begin(), H.begin() + padded_n);
        out.ipp = ipp_create(tr, Q, Gf, Hf, std::move(Gv), std::move(Hv), std::move(l_vec), std::move(r_vec));
        return R1CS_OK;
    }
};

struct verifier : constraint_system {
    transcript &tr;
    std::vector<lincomb> constraints;
    size_t num_vars;
    std::vector<bytes32> V;

    explicit verifier(transcript &t) : tr(t), num_vars(0) { tr.r1cs_domain_sep(); }

    variable commit(const bytes32 &commitment) {
        uint32_t i = (uint32_t)V.size();
        V.push_back(commitment);
        tr.append_point("V", commitment.data());
        return variable{VAR_COMMITTED, i};
    }
    void multiply(lincomb left, lincomb right, variable &l, variable &r, variable &o) override {
        uint32_t i = (uint32_t)num_vars++;
        l = variable{VAR_MUL_LEFT, i};
        r = variable{VAR_MUL_RIGHT, i};
        o = variable{VAR_MUL_OUT, i};
        left.terms.push_back({l, sc_neg(sc_one())});
        right.terms.push_back({r, sc_neg(sc_one())});
        constrain(std::move(left));
        constrain(std::move(right));
    }
    void constrain(lincomb lc) override { constraints.push_back(std::move(lc)); }

    void flattened_constraints(const sc &z, std::vector<sc> &wL, std::vector<sc> &wR, std::vector<sc> &wO, std::vector<sc> &wV, sc &wc) const {
        size_t n = num_vars, m = V.size();
        wL.assign(n, sc_zero()); wR.assign(n, sc_zero()); wO.assign(n, sc_zero()); wV.assign(m, sc_zero());
        wc = sc_zero();
        sc exp_z = z;
        for (auto &lc : constraints) {
            for (auto &t : lc.terms) {
                sc term = sc_mul(exp_z, t.second);
                switch (t.first.kind) {
                    case VAR_MUL_LEFT: wL[t.first.idx] = sc_add(wL[t.first.idx], term); break;
                    case VAR_MUL_RIGHT: wR[t.first.idx] = sc_add(wR[t.first.idx], term); break;
                    case VAR_MUL_OUT: wO[t.first.idx] = sc_add(wO[t.first.idx], term); break;
                    case VAR_COMMITTED: wV[t.first.idx] = sc_sub(wV[t.first.idx], term); break;
                    case VAR_ONE: wc = sc_sub(wc, term); break;
                }
            }
            exp_z = sc_mul(exp_z, z);
        }
    }

    // Verifier::verify (verify.rs:88). If mega_scalars/mega_points are given, the assembled mega-check is
    // also exported (used by tests that compare the product's GPU-side scalar assembly).
    int verify(const r1cs_proof &proof, const pedersen_gens &pc, const bulletproof_gens &bp, const uint8_t external_rng32[32],
               int threads = 1, std::vector<sc> *mega_scalars = nullptr) {
        tr.append_u64("m", V.size());
        size_t n1 = num_vars;
        if (!tr.validate_and_append_point("A_I1", proof.A_I1.data())) return R1CS_VERIFICATION_ERROR;
        if (!tr.validate_and_append_point("A_O1", proof.A_O1.data())) return R1CS_VERIFICATION_ERROR;
        if (!tr.validate_and_append_point("S1", proof.S1.data())) return R1CS_VERIFICATION_ERROR;
        tr.r1cs_1phase_domain_sep();
        size_t n = num_vars;
        size_t n2 = n - n1;
        size_t padded_n = next_pow2(n);
        size_t pad = padded_n - n;
        if (bp.gens_capacity < padded_n) return R1CS_INVALID_GENERATORS_LENGTH;
        const std::vector<ge> &G = bp.G[0], &H = bp.H[0];

        tr.append_point("A_I2", proof.A_I2.data());
        tr.append_point("A_O2", proof.A_O2.data());
        tr.append_point("S2", proof.S2.data());
        sc y = tr.challenge_scalar("y");
        sc z = tr.challenge_scalar("z");
        if (!tr.validate_and_append_point("T_1", proof.T_1.data())) return R1CS_VERIFICATION_ERROR;
        if (!tr.validate_and_append_point("T_3", proof.T_3.data())) return R1CS_VERIFICATION_ERROR;
        if (!tr.validate_and_append_point("T_4", proof.T_4.data())) return R1CS_VERIFICATION_ERROR;
        if (!tr.validate_and_append_point("T_5", proof.T_5.data())) return R1CS_VERIFICATION_ERROR;
        if (!tr.validate_and_append_point("T_6", proof.T_6.data())) return R1CS_VERIFICATION_ERROR;
        sc u = tr.challenge_scalar("u");
        sc x = tr.challenge_scalar("x");
        tr.append_scalar("t_x", proof.t_x);
        tr.append_scalar("t_x_blinding", proof.t_x_blinding);
        tr.append_scalar("e_blinding", proof.e_blinding);
        sc w = tr.challenge_scalar("w");

        std::vector<sc> wL, wR, wO, wV;
        sc wc;
        flattened_constraints(z, wL, wR, wO, wV, wc);

        std::vector<sc> u_sq, u_inv_sq, s;
        if (!ipp_verification_scalars(proof.ipp, padded_n, tr, u_sq, u_inv_sq, s)) return R1CS_VERIFICATION_ERROR;
        sc a = proof.ipp.a, b = proof.ipp.b;

        sc y_inv = sc_invert(y);
        std::vector<sc> y_inv_vec(padded_n);
        {
            sc e = sc_one();
            for (size_t i = 0; i < padded_n; i++) { y_inv_vec[i] = e; e = sc_mul(e, y_inv); }
        }
        std::vector<sc> yneg_wR(padded_n, sc_zero());
        for (size_t i = 0; i < n; i++) yneg_wR[i] = sc_mul(wR[i], y_inv_vec[i]);
        sc delta = sc_inner_product(yneg_wR.data(), wL.data(), n);

        std::vector<sc> g_scalars(padded_n), h_scalars(padded_n);
        sc minus_one = sc_neg(sc_one());
        for (size_t i = 0; i < padded_n; i++) {
            sc u_or_1 = (i < n1) ? sc_one() : u;
            g_scalars[i] = sc_mul(u_or_1, sc_sub(sc_mul(x, yneg_wR[i]), sc_mul(a, s[i])));
            sc wLi = (i < n) ? wL[i] : sc_zero();
            sc wOi = (i < n) ? wO[i] : sc_zero();
            sc inner = sc_sub(sc_add(sc_mul(x, wLi), wOi), sc_mul(b, s[padded_n - 1 - i]));
            h_scalars[i] = sc_mul(u_or_1, sc_add(sc_mul(y_inv_vec[i], inner), minus_one));
        }
        (void)n2; (void)pad;

        transcript_rng rng = tr.build_rng().finalize(external_rng32);
        sc r = rng.random_scalar();

        sc xx = sc_mul(x, x), rxx = sc_mul(r, xx), xxx = sc_mul(x, xx);
        std::vector<sc> ms;
        std::vector<bytes32> mp_c;
        std::vector<ge> mp;
        auto push_c = [&](const sc &sv, const bytes32 &pt) { ms.push_back(sv); mp_c.push_back(pt); };
        push_c(x, proof.A_I1); push_c(xx, proof.A_O1); push_c(xxx, proof.S1);
        push_c(sc_mul(u, x), proof.A_I2); push_c(sc_mul(u, xx), proof.A_O2); push_c(sc_mul(u, xxx), proof.S2);
        for (size_t i = 0; i < V.size(); i++) push_c(sc_mul(wV[i], rxx), V[i]);
        push_c(sc_mul(r, x), proof.T_1);
        push_c(sc_mul(rxx, x), proof.T_3);
        push_c(sc_mul(rxx, xx), proof.T_4);
        push_c(sc_mul(rxx, xxx), proof.T_5);
        push_c(sc_mul(sc_mul(rxx, xx), xx), proof.T_6);
        // decompress the first group; any failure => VerificationError (optional_multiscalar_mul -> None)
        bool ok = true;
        for (auto &c : mp_c) {
            ge p;
            if (!ge_decompress(p, c.data())) { ok = false; break; }
            mp.push_back(p);
        }
        sc b_scalar = sc_add(sc_mul(w, sc_sub(proof.t_x, sc_mul(a, b))),
                             sc_mul(r, sc_sub(sc_mul(xx, sc_add(wc, delta)), proof.t_x)));
        sc bb_scalar = sc_sub(sc_neg(proof.e_blinding), sc_mul(r, proof.t_x_blinding));
        ms.push_back(b_scalar); if (ok) mp.push_back(pc.B);
        ms.push_back(bb_scalar); if (ok) mp.push_back(pc.B_blinding);
        for (size_t i = 0; i < padded_n; i++) { ms.push_back(g_scalars[i]); if (ok) mp.push_back(G[i]); }
        for (size_t i = 0; i < padded_n; i++) { ms.push_back(h_scalars[i]); if (ok) mp.push_back(H[i]); }
        for (size_t i = 0; i < u_sq.size(); i++) {
            ms.push_back(u_sq[i]);
            ge p;
            if (ok && !ge_decompress(p, proof.ipp.L[i].data())) ok = false;
            if (ok) mp.push_back(p);
        }
        for (size_t i = 0; i < u_inv_sq.size(); i++) {
            ms.push_back(u_inv_sq[i]);
            ge p;
            if (ok && !ge_decompress(p, proof.ipp.R[i].data())) ok = false;
            if (ok) mp.push_back(p);
        }
        if (mega_scalars) *mega_scalars = ms;
        if (!ok) return R1CS_VERIFICATION_ERROR;
        ge mega = msm_pippenger(ms.data(), mp.data(), ms.size(), threads);
        if (!ge_is_identity(mega)) return R1CS_VERIFICATION_ERROR;
        return R1CS_OK;
    }
};

}  // namespace orc
