// Host side of the blind-bid R1CS circuit: the gadgets of src/gadgets.rs written once against an abstract
// constraint-system interface and instantiated twice —
//   * `recorder`  builds the circuit TEMPLATE for a given (number of commitments, number of toggles): the sparse
//     structure of all constraints in a GPU-friendly CSR form (variable terms carry only a sign, constant terms carry a
//     symbolic reference into a small per-proof table of public values). The structure depends only on the list length
//     L, so it is built once, uploaded once, and shared by every proof in a batch.
//   * `evaluator` computes the prover's witness (a_L, a_R, a_O) for one bid: linear combinations collapse to their
//     values, `multiply` is one mod-l product.
// This replaces the bulletproofs `ConstraintSystem` / `LinearCombination` / `Variable` machinery the reference drives at
// src/blindbid/proof.rs:50-88 and src/blindbid/verify.rs:51-88 (bulletproofs 1.0.4 @ 4a05305).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <vector>
#include "keccak.h"
#include "sc25519.cuh"

namespace bbp {

static const uint32_t MIMC_ROUNDS = 90;   // src/gadgets.rs:4

// src/blindbid/mod.rs:7-24: c_0 = wide_reduce(SHA512("blind bid")), c_{i+1} = wide_reduce(SHA512(bytes(c_i)))
inline const std::vector<sc> &mimc_constants() {
    static const std::vector<sc> C = [] {
        std::vector<sc> c;
        uint8_t h[64];
        sha512(h, "blind bid", 9);
        for (uint32_t i = 0; i < MIMC_ROUNDS; i++) {
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

// ---- variables and symbolic constants ----------------------------------------------------------------------------
enum var_kind : uint32_t { VK_COMMITTED = 0, VK_LEFT = 1, VK_RIGHT = 2, VK_OUT = 3 };
// per-proof public value table: [0] = 1, [1..91) = MiMC constants, [91] = seed, [92] = q (score), [93] = z_img, [94+i] = item i
enum : uint32_t { PV_ONE = 0, PV_MIMC = 1, PV_SEED = 91, PV_Q = 92, PV_ZIMG = 93, PV_ITEM = 94 };

// ---- recorder: symbolic linear combinations ------------------------------------------------------------------------
struct sym_term {
    uint32_t ref;    // variable: kind << 28 | index ; constant: public value index
    bool is_const;
    bool neg;
};
struct sym_lc {
    std::vector<sym_term> t;
};

struct circuit_template {
    uint32_t n_commit = 0, n_toggle = 0;
    uint32_t n1 = 0;          // multipliers
    uint32_t q = 0;           // constraints
    uint32_t m = 0;           // commitments
    uint32_t n_pub = 0;       // length of the public value table
    uint32_t n_pub_shared = 0;   // its first entries that are the same for every proof of the circuit (blind bid: 1 + MiMC constants)
    // CSR over target rows: rows [0,n1) = wL, [n1,2n1) = wR, [2n1,3n1) = wO, [3n1,3n1+m) = wV.
    // entry = constraint index j | sign bit (bit 31 set: subtract z^(j+1))
    std::vector<uint32_t> row_ptr, entries;
    // generic circuits only (empty for the blind-bid template, whose variable coefficients are all +-1): per entry the
    // index of the term's coefficient in the public value table (0 = one), and that table itself ([0] = 1, then the
    // circuit's coefficients and constants: it belongs to the circuit, not to the proof)
    std::vector<uint32_t> coef;
    std::vector<sc> coef_table;
    // constant terms (verifier only): wc = - sum sign * z^(j+1) * pub[idx];  entry j | sign<<31, paired with idx
    std::vector<uint32_t> const_j, const_idx;
};

struct recorder {
    typedef sym_lc LC;
    circuit_template *tpl;
    std::vector<std::vector<uint32_t>> rows;   // filled per target row, flattened at the end
    uint32_t n_mul = 0, n_con = 0;

    explicit recorder(circuit_template *t) : tpl(t) {}

    LC constant(uint32_t pv_index) const { LC l; l.t.push_back({pv_index, true, false}); return l; }
    LC zero() const { return LC(); }   // Scalar::zero().into(): a constant term with coefficient 0 contributes nothing
    LC var(var_kind k, uint32_t i) const { LC l; l.t.push_back({((uint32_t)k << 28) | i, false, false}); return l; }
    LC add(LC a, const LC &b) const { a.t.insert(a.t.end(), b.t.begin(), b.t.end()); return a; }
    LC sub(LC a, const LC &b) const {
        for (auto x : b.t) { x.neg = !x.neg; a.t.push_back(x); }
        return a;
    }
    // Terms that cancel inside one constraint (+t and -t of the same variable or constant) are dropped before they reach
    // the CSR: the toggle-sum constraints of one_of_many_gadget (src/gadgets.rs:119) are sum[i-1] + t_i - sum[i] with
    // sum[i] = sum[i-1] + t_i, i.e. 2i + 1 pairs that all cancel — O(L^2) entries that contribute exactly zero to every
    // flattened weight. The constraint still takes its index j (its z power), so nothing else moves.
    static void cancel(const LC &lc, std::vector<sym_term> &out) {
        std::map<uint64_t, int32_t> net;   // (is_const, ref) -> signed multiplicity
        for (auto &x : lc.t) net[((uint64_t)x.is_const << 32) | x.ref] += x.neg ? -1 : 1;
        for (auto &x : lc.t) {
            auto it = net.find(((uint64_t)x.is_const << 32) | x.ref);
            if (it->second == 0) continue;
            bool neg = it->second < 0;
            for (int32_t k = 0; k < (neg ? -it->second : it->second); k++) out.push_back({x.ref, x.is_const, neg});
            it->second = 0;
        }
    }
    void constrain(const LC &lc_in) {
        uint32_t j = n_con++;
        sym_lc lc;
        cancel(lc_in, lc.t);
        for (auto &x : lc.t) {
            if (x.is_const) {
                tpl->const_j.push_back(j | (x.neg ? 0x80000000u : 0u));
                tpl->const_idx.push_back(x.ref);
                continue;
            }
            uint32_t kind = x.ref >> 28, idx = x.ref & 0x0fffffffu;
            size_t row;
            bool neg = x.neg;
            switch (kind) {
                case VK_LEFT: row = idx; break;
                case VK_RIGHT: row = (size_t)n_mul_cap + idx; break;
                case VK_OUT: row = 2 * (size_t)n_mul_cap + idx; break;
                default: row = 3 * (size_t)n_mul_cap + idx; neg = !neg; break;   // wV -= z^(j+1) * c
            }
            if (rows.size() <= row) rows.resize(row + 1);
            rows[row].push_back(j | (neg ? 0x80000000u : 0u));
        }
    }
    // multiply: allocates (L_i, R_i, O_i) and adds  left - L_i = 0,  right - R_i = 0
    void multiply(const LC &left, const LC &right, LC &l, LC &r, LC &o) {
        uint32_t i = n_mul++;
        l = var(VK_LEFT, i); r = var(VK_RIGHT, i); o = var(VK_OUT, i);
        constrain(sub(left, l));
        constrain(sub(right, r));
    }
    uint32_t n_mul_cap = 0;   // rows are addressed with a fixed stride so that they can be filled before n1 is known
};

// ---- evaluator: the prover's witness ---------------------------------------------------------------------------------
struct evaluator {
    typedef sc LC;
    const sc *pub;                 // public value table of this proof
    const sc *committed;           // values behind the commitments
    sc *a_L, *a_R, *a_O;           // witness rows, written in place (capacity: the template's n1)
    uint32_t count = 0;

    LC constant(uint32_t pv_index) const { return pub[pv_index]; }
    LC zero() const { return sc_zero(); }
    LC add(const LC &a, const LC &b) const { return sc_add(a, b); }
    LC sub(const LC &a, const LC &b) const { return sc_sub(a, b); }
    void constrain(const LC &) {}
    void multiply(const LC &left, const LC &right, LC &l, LC &r, LC &o) {
        l = left; r = right; o = sc_mul(left, right);
        a_L[count] = l; a_R[count] = r; a_O[count] = o;
        count++;
    }
};

// ---- gadgets (src/gadgets.rs), generic over the two instantiations ---------------------------------------------------
// src/gadgets.rs:37-68
template <class CS>
typename CS::LC mimc_gadget(CS &cs, const typename CS::LC &left, const typename CS::LC &right) {
    typedef typename CS::LC LC;
    LC x = left;
    LC l, r, a2, a3, a4, a7;
    for (uint32_t i = 0; i < MIMC_ROUNDS; i++) {
        LC a = cs.add(cs.add(x, right), cs.constant(PV_MIMC + i));
        cs.multiply(a, a, l, r, a2);
        cs.multiply(a2, a, l, r, a3);
        cs.multiply(a2, a2, l, r, a4);
        cs.multiply(a4, a3, l, r, a7);
        x = a7;
    }
    return cs.add(x, right);
}

// src/gadgets.rs:134-140
template <class CS>
void boolean_gadget(CS &cs, const typename CS::LC &a) {
    typename CS::LC l, r, c;
    cs.multiply(a, cs.sub(cs.constant(PV_ONE), a), l, r, c);
    cs.constrain(c);
}

// src/gadgets.rs:88-132 (the reference indexes toggle[0]: callers guarantee at least one toggle)
template <class CS>
void one_of_many_gadget(CS &cs, const typename CS::LC &x, const std::vector<typename CS::LC> &toggle, const std::vector<typename CS::LC> &items) {
    typedef typename CS::LC LC;
    size_t n = toggle.size();
    for (size_t i = 0; i < n; i++) boolean_gadget(cs, toggle[i]);
    std::vector<LC> sum;
    sum.push_back(toggle[0]);
    for (size_t i = 1; i < n; i++) sum.push_back(cs.add(sum[i - 1], toggle[i]));
    for (size_t i = 1; i < n; i++) cs.constrain(cs.sub(cs.add(sum[i - 1], toggle[i]), sum[i]));
    cs.constrain(cs.sub(sum[n - 1], cs.constant(PV_ONE)));
    for (size_t i = 0; i < n; i++) {
        LC l, r, left, right;
        cs.multiply(items[i], toggle[i], l, r, left);
        cs.multiply(toggle[i], x, l, r, right);
        cs.constrain(cs.sub(left, right));
    }
}

// src/gadgets.rs:70-86
template <class CS>
void score_gadget(CS &cs, const typename CS::LC &d, const typename CS::LC &y, const typename CS::LC &y_inv, const typename CS::LC &q) {
    typename CS::LC l, r, one_var, q_var;
    cs.multiply(y, y_inv, l, r, one_var);
    cs.constrain(cs.sub(one_var, cs.constant(PV_ONE)));
    cs.multiply(d, y_inv, l, r, q_var);
    cs.constrain(cs.sub(q, q_var));
}

// src/gadgets.rs:6-34
template <class CS>
void proof_gadget(CS &cs, const typename CS::LC &d, const typename CS::LC &k, const typename CS::LC &y_inv, const typename CS::LC &q,
                  const typename CS::LC &z_img, const typename CS::LC &seed, const std::vector<typename CS::LC> &toggle,
                  const std::vector<typename CS::LC> &items) {
    typedef typename CS::LC LC;
    LC m = mimc_gadget(cs, k, cs.zero());
    LC x = mimc_gadget(cs, d, m);
    one_of_many_gadget(cs, x, toggle, items);
    LC y = mimc_gadget(cs, seed, x);
    LC z = mimc_gadget(cs, seed, m);
    cs.constrain(cs.sub(z_img, z));
    score_gadget(cs, d, y, y_inv, q);
}

// ---- template cache -------------------------------------------------------------------------------------------------
// Wiring as in src/blindbid/proof.rs:74-85 / verify.rs:74-85: d = V[0], k = V[1], y_inv = V[3] (V[2], the commitment to
// y, is committed but never wired), toggles = V[n_commit ..], items = the first n_toggle public list entries.
// shape of the circuit without building it: n1 = 4 MiMC x 90 rounds x 4 + L booleans + 2 L membership + 2 score multipliers,
// q = 2 n1 + 1 + 2 + (L - 1) + 1 + L + L constraints, m = commitments + toggles
inline uint32_t blindbid_n1(uint32_t n_toggle) { return 16 * MIMC_ROUNDS + 3 * n_toggle + 2; }
inline uint32_t blindbid_q(uint32_t n_toggle) { return 2 * blindbid_n1(n_toggle) + 3 + 3 * n_toggle; }
static const uint32_t BLINDBID_MAX_TOGGLES = 1u << 16;      // the recorder's toggle-sum LCs are O(L^2) terms while being built
static const uint32_t BLINDBID_MAX_COMMITMENTS = 1u << 16;
static const size_t TEMPLATE_CACHE_MAX = 64;                // (n_commit, n_toggle) comes from untrusted requests: bounded cache

inline std::shared_ptr<const circuit_template> blindbid_template(uint32_t n_commit, uint32_t n_toggle) {
    static std::mutex mu;
    static std::map<uint64_t, std::shared_ptr<const circuit_template>> cache;
    std::lock_guard<std::mutex> lock(mu);
    uint64_t key = ((uint64_t)n_commit << 32) | n_toggle;
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    if (cache.size() >= TEMPLATE_CACHE_MAX) cache.clear();   // shared_ptr keeps templates in use alive
    auto tpl = std::make_shared<circuit_template>();
    tpl->n_commit = n_commit; tpl->n_toggle = n_toggle;
    tpl->m = n_commit + n_toggle;
    tpl->n_pub = PV_ITEM + n_toggle;
    tpl->n_pub_shared = PV_SEED;
    recorder rec(tpl.get());
    rec.n_mul_cap = 4 * 4 * MIMC_ROUNDS + 3 * n_toggle + 2;
    std::vector<sym_lc> toggles, items;
    for (uint32_t i = 0; i < n_toggle; i++) {
        toggles.push_back(rec.var(VK_COMMITTED, n_commit + i));
        items.push_back(rec.constant(PV_ITEM + i));
    }
    proof_gadget(rec, rec.var(VK_COMMITTED, 0), rec.var(VK_COMMITTED, 1), rec.var(VK_COMMITTED, 3), rec.constant(PV_Q), rec.constant(PV_ZIMG),
                 rec.constant(PV_SEED), toggles, items);
    tpl->n1 = rec.n_mul;
    tpl->q = rec.n_con;
    // flatten rows (stride n_mul_cap) into the compact CSR (stride n1)
    uint32_t n1 = tpl->n1, cap = rec.n_mul_cap;
    size_t n_rows = 3 * (size_t)n1 + tpl->m;
    tpl->row_ptr.assign(n_rows + 1, 0);
    auto src_row = [&](size_t r) -> size_t {
        if (r < 3 * (size_t)n1) return (r / n1) * cap + (r % n1);
        return 3 * (size_t)cap + (r - 3 * (size_t)n1);
    };
    for (size_t r = 0; r < n_rows; r++) {
        size_t s = src_row(r);
        if (s < rec.rows.size())
            for (uint32_t e : rec.rows[s]) tpl->entries.push_back(e);
        tpl->row_ptr[r + 1] = (uint32_t)tpl->entries.size();
    }
    cache[key] = tpl;
    return tpl;
}

// ---- generic circuits: the flattened form of a bulletproofs ConstraintSystem ---------------------------------------
// What `ConstraintSystem::{multiply, constrain}` (src/gadgets.rs:30,53 and every other call site) leave behind is a list
// of linear constraints over (Variable, Scalar) terms; `multiply` additionally allocates one multiplier and contributes
// its two constraints  left - L_i = 0,  right - R_i = 0  (bulletproofs 1.0.4 r1cs/prover.rs, verifier.rs). A caller
// that recorded its circuit hands the constraints over in CSR form: constraint j owns terms con_ptr[j] .. con_ptr[j+1],
// term t = (term_var[t] = kind << 28 | index, term_coeff[32 t ..] = canonical scalar), kinds as var_kind plus VK_ONE.
enum : uint32_t { VK_ONE = 4 };
// returns nullptr on malformed input (index out of range, non-canonical coefficient, unknown kind)
inline std::shared_ptr<const circuit_template> generic_template(uint32_t n_mul, uint32_t n_commit, uint32_t n_con, const uint32_t *con_ptr,
                                                                const uint32_t *term_var, const uint8_t *term_coeff) {
    auto tpl = std::make_shared<circuit_template>();
    tpl->n_commit = n_commit; tpl->n_toggle = 0; tpl->m = n_commit; tpl->n1 = n_mul; tpl->q = n_con;
    tpl->coef_table.push_back(sc_one());
    const sc one = sc_one(), minus_one = sc_neg(sc_one());
    std::map<std::vector<uint8_t>, uint32_t> known;   // coefficient bytes -> table index
    auto table_index = [&](const sc &c, const uint8_t *bytes) -> uint32_t {
        if (sc_eq(c, one)) return 0;
        std::vector<uint8_t> key(bytes, bytes + 32);
        auto it = known.find(key);
        if (it != known.end()) return it->second;
        uint32_t idx = (uint32_t)tpl->coef_table.size();
        tpl->coef_table.push_back(c);
        known.emplace(std::move(key), idx);
        return idx;
    };
    const size_t n_rows = 3 * (size_t)n_mul + n_commit;
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> rows(n_rows);   // (entry, coefficient index)
    if (n_con && con_ptr[0] != 0) return nullptr;
    bool any_coef = false;
    for (uint32_t j = 0; j < n_con; j++) {
        if (con_ptr[j + 1] < con_ptr[j]) return nullptr;
        for (uint32_t t = con_ptr[j]; t < con_ptr[j + 1]; t++) {
            const uint32_t kind = term_var[t] >> 28, idx = term_var[t] & 0x0fffffffu;
            sc c;
            if (!sc_from_canonical(c, term_coeff + 32 * (size_t)t)) return nullptr;
            if (sc_iszero(c)) continue;   // contributes nothing to any flattened weight
            if (kind == VK_ONE) {
                tpl->const_j.push_back(j);
                tpl->const_idx.push_back(table_index(c, term_coeff + 32 * (size_t)t));
                continue;
            }
            size_t row;
            bool neg = false;
            switch (kind) {
                case VK_LEFT: if (idx >= n_mul) return nullptr; row = idx; break;
                case VK_RIGHT: if (idx >= n_mul) return nullptr; row = (size_t)n_mul + idx; break;
                case VK_OUT: if (idx >= n_mul) return nullptr; row = 2 * (size_t)n_mul + idx; break;
                case VK_COMMITTED: if (idx >= n_commit) return nullptr; row = 3 * (size_t)n_mul + idx; neg = true; break;   // wV -= z^(j+1) c
                default: return nullptr;
            }
            uint32_t ci = 0;
            if (sc_eq(c, minus_one)) neg = !neg;
            else ci = table_index(c, term_coeff + 32 * (size_t)t);
            any_coef = any_coef || ci != 0;
            rows[row].push_back({j | (neg ? 0x80000000u : 0u), ci});
        }
    }
    tpl->n_pub = (uint32_t)tpl->coef_table.size();
    tpl->row_ptr.assign(n_rows + 1, 0);
    for (size_t r = 0; r < n_rows; r++) {
        for (auto &e : rows[r]) { tpl->entries.push_back(e.first); tpl->coef.push_back(e.second); }
        tpl->row_ptr[r + 1] = (uint32_t)tpl->entries.size();
    }
    if (!any_coef) tpl->coef.clear();
    return tpl;
}

// native MiMC-x^7 (what the gadget constrains); used for synthetic bids / client-side helpers
inline sc mimc_hash(const sc &left, const sc &right) {
    const std::vector<sc> &c = mimc_constants();
    sc x = left;
    for (uint32_t i = 0; i < MIMC_ROUNDS; i++) {
        sc a = sc_add(sc_add(x, right), c[i]);
        sc a2 = sc_mul(a, a), a3 = sc_mul(a2, a), a4 = sc_mul(a2, a2);
        x = sc_mul(a4, a3);
    }
    return sc_add(x, right);
}

// fills the public value table of one proof (length PV_ITEM + n_items)
inline void fill_public_values(std::vector<sc> &pub, const sc &seed, const sc &q, const sc &z_img, const sc *items, uint32_t n_items) {
    const std::vector<sc> &c = mimc_constants();
    pub.resize(PV_ITEM + n_items);
    pub[PV_ONE] = sc_one();
    for (uint32_t i = 0; i < MIMC_ROUNDS; i++) pub[PV_MIMC + i] = c[i];
    pub[PV_SEED] = seed; pub[PV_Q] = q; pub[PV_ZIMG] = z_img;
    for (uint32_t i = 0; i < n_items; i++) pub[PV_ITEM + i] = items[i];
}

}  // namespace bbp
