// ORACLE — test infrastructure only (see fe.h header).
// The reference's own code, restated: MiMC constants (src/blindbid/mod.rs:7-24), the circuit gadgets
// (src/gadgets.rs:6-140) and the prove / verify drivers (src/blindbid/proof.rs:36-91,
// src/blindbid/verify.rs:47-89), plus the native MiMC helper used to make synthetic bids
// (SURVEY.md §8d config 3, §8f-4).
#pragma once
#include "r1cs.h"

namespace orc {

static const size_t MIMC_ROUNDS = 90;   // src/gadgets.rs:4

// src/blindbid/mod.rs:7-24: c0 = wide_reduce(SHA512("blind bid")), c_{i+1} = wide_reduce(SHA512(c_i))
static inline const std::vector<sc> &mimc_constants() {
    static const std::vector<sc> C = [] {
        std::vector<sc> c;
        uint8_t h[64];
        sha512(h, (const uint8_t *)"blind bid", 9);
        for (size_t i = 0; i < MIMC_ROUNDS; i++) {
            sc k = sc_from_wide(h);
            c.push_back(k);
            uint8_t kb[32];
            sc_tobytes(kb, k);
            sha512(h, kb, 32);
        }
        return c;
    }();
    return C;
}

// native MiMC-x^7 (the function the gadget at src/gadgets.rs:37-68 constrains)
static inline sc mimc_hash(const sc &left, const sc &right) {
    const std::vector<sc> &c = mimc_constants();
    sc x = left;
    for (size_t i = 0; i < MIMC_ROUNDS; i++) {
        sc a = sc_add(sc_add(x, right), c[i]);
        sc a2 = sc_mul(a, a), a3 = sc_mul(a2, a), a4 = sc_mul(a2, a2);
        x = sc_mul(a4, a3);
    }
    return sc_add(x, right);
}

// src/gadgets.rs:37-68
static inline lincomb mimc_gadget(constraint_system &cs, const lincomb &left, const lincomb &right, const std::vector<sc> &constants) {
    lincomb x = left;
    const lincomb &key = right;
    variable l, r, a_2, a_3, a_4, a_7;
    for (size_t i = 0; i < MIMC_ROUNDS; i++) {
        lincomb a = x + key + lincomb(constants[i]);
        cs.multiply(a, a, l, r, a_2);
        cs.multiply(lincomb(a_2), a, l, r, a_3);
        cs.multiply(lincomb(a_2), lincomb(a_2), l, r, a_4);
        cs.multiply(lincomb(a_4), lincomb(a_3), l, r, a_7);
        x = lincomb(a_7);
    }
    return x + key;
}

// src/gadgets.rs:70-86
static inline void score_gadget(constraint_system &cs, const lincomb &d, const lincomb &y, const lincomb &y_inv, const lincomb &q) {
    variable l, r, one_var, q_var;
    cs.multiply(y, y_inv, l, r, one_var);
    cs.constrain(lincomb(one_var) - lincomb(sc_one()));
    cs.multiply(d, y_inv, l, r, q_var);
    cs.constrain(q - lincomb(q_var));
}

// src/gadgets.rs:134-140
static inline void boolean_gadget(constraint_system &cs, const lincomb &a1) {
    variable l, r, c_var;
    cs.multiply(a1, lincomb(sc_one()) - a1, l, r, c_var);
    cs.constrain(lincomb(c_var));
}

// src/gadgets.rs:88-132. The reference indexes toggle[0] unconditionally (panics on an empty list);
// callers of the oracle must pass at least one toggle.
static inline void one_of_many_gadget(constraint_system &cs, const lincomb &x, const std::vector<variable> &toggle, const std::vector<lincomb> &items) {
    size_t n = toggle.size();
    for (size_t i = 0; i < n; i++) boolean_gadget(cs, lincomb(toggle[i]));
    std::vector<lincomb> toggle_sum;
    toggle_sum.push_back(lincomb(toggle[0]));
    for (size_t i = 1; i < n; i++) toggle_sum.push_back(toggle_sum[i - 1] + lincomb(toggle[i]));
    for (size_t i = 1; i < n; i++) {
        lincomb prev = toggle_sum[i - 1];
        lincomb cur_toggle(toggle[i]);
        lincomb cur_sum = toggle_sum[i];
        toggle_sum[i] = toggle_sum[i - 1] + lincomb(toggle[i]);
        cs.constrain(prev + cur_toggle - cur_sum);
    }
    cs.constrain(toggle_sum[n - 1] - lincomb(sc_one()));
    for (size_t i = 0; i < n; i++) {
        variable l, r, left, right;
        cs.multiply(items[i], lincomb(toggle[i]), l, r, left);
        cs.multiply(lincomb(toggle[i]), x, l, r, right);
        cs.constrain(lincomb(left) - lincomb(right));
    }
}

// src/gadgets.rs:6-34
static inline void proof_gadget(constraint_system &cs, const lincomb &d, const lincomb &k, const lincomb &y_inv, const lincomb &q,
                                const lincomb &z_img, const lincomb &seed, const std::vector<sc> &constants,
                                const std::vector<variable> &toggle, const std::vector<lincomb> &items) {
    lincomb m = mimc_gadget(cs, k, lincomb(sc_zero()), constants);
    lincomb x = mimc_gadget(cs, d, m, constants);
    one_of_many_gadget(cs, x, toggle, items);
    lincomb y = mimc_gadget(cs, seed, x, constants);
    lincomb z = mimc_gadget(cs, seed, m, constants);
    cs.constrain(z_img - z);
    score_gadget(cs, d, y, y_inv, q);
}

struct blindbid_gens {   // generate_cs_transcript() minus the transcript (src/blindbid/mod.rs:34-40), cached
    pedersen_gens pc;
    bulletproof_gens bp;
    blindbid_gens() : pc(), bp(2048, 1) {}
};
static inline const blindbid_gens &blindbid_generators() {
    static const blindbid_gens g;
    return g;
}

struct blindbid_proof {
    r1cs_proof proof;
    std::vector<bytes32> commitments, t_c;
};

// Proof::prove (src/blindbid/proof.rs:36-91). blindings: 4 + L scalars replacing Scalar::random(thread_rng)
// at :57,:64; external_rng32 replaces the thread_rng bytes inside Prover::prove.
static inline int blindbid_prove(const sc &d, const sc &k, const sc &y, const sc &y_inv, const sc &q, const sc &z_img, const sc &seed,
                                 const std::vector<sc> &pub_list, uint64_t toggle, const std::vector<sc> &blindings,
                                 const uint8_t external_rng32[32], blindbid_proof &out) {
    const blindbid_gens &g = blindbid_generators();
    transcript tr("BlindBidProofGadget");
    prover pr(g.pc, tr);
    const sc vals[4] = {d, k, y, y_inv};
    std::vector<variable> vars;
    out.commitments.resize(4);
    for (int i = 0; i < 4; i++) vars.push_back(pr.commit(vals[i], blindings[i], out.commitments[i]));
    std::vector<variable> t_v;
    out.t_c.resize(pub_list.size());
    for (size_t i = 0; i < pub_list.size(); i++)
        t_v.push_back(pr.commit(sc_from_u64((uint64_t)i == toggle ? 1 : 0), blindings[4 + i], out.t_c[i]));
    std::vector<lincomb> l_v;
    for (auto &b : pub_list) l_v.push_back(lincomb(b));
    proof_gadget(pr, lincomb(vars[0]), lincomb(vars[1]), lincomb(vars[3]), lincomb(q), lincomb(z_img), lincomb(seed),
                 mimc_constants(), t_v, l_v);
    return pr.prove(g.bp, external_rng32, out.proof);
}

// Verify::verify (src/blindbid/verify.rs:47-89). The reference indexes vars[0], vars[1], vars[3] and so
// panics with fewer than 4 commitments; the oracle reports that as a format error.
static inline int blindbid_verify(const blindbid_proof &p, const sc &score, const sc &z_img, const sc &seed, const std::vector<sc> &pub_list,
                                  const uint8_t external_rng32[32], int threads = 1, std::vector<sc> *mega_scalars = nullptr) {
    if (p.commitments.size() < 4 || p.t_c.empty() || pub_list.size() < p.t_c.size()) return R1CS_FORMAT_ERROR;
    const blindbid_gens &g = blindbid_generators();
    transcript tr("BlindBidProofGadget");
    verifier ve(tr);
    std::vector<variable> vars, t_c_v;
    for (auto &c : p.commitments) vars.push_back(ve.commit(c));
    for (auto &c : p.t_c) t_c_v.push_back(ve.commit(c));
    std::vector<lincomb> l_v;
    for (auto &b : pub_list) l_v.push_back(lincomb(b));
    proof_gadget(ve, lincomb(vars[0]), lincomb(vars[1]), lincomb(vars[3]), lincomb(score), lincomb(z_img), lincomb(seed),
                 mimc_constants(), t_c_v, l_v);
    return ve.verify(p.proof, g.pc, g.bp, external_rng32, threads, mega_scalars);
}

}  // namespace orc
