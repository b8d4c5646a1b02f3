// ORACLE — test infrastructure only (see fe.h header).
// bulletproofs@4a05305 InnerProductProof::{create, verification_scalars, to_bytes, from_bytes}
// (SURVEY.md §2.2 U7, §8 a-6, a-8). The reference reaches it through Prover::prove
// (src/blindbid/proof.rs:88) and Verifier::verify (src/blindbid/verify.rs:88).
#pragma once
#include "ge.h"
#include "merlin.h"
#include "msm.h"
#include <array>
#include <vector>

namespace orc {

typedef std::array<uint8_t, 32> bytes32;

struct ipp_proof {
    std::vector<bytes32> L, R;
    sc a, b;
};

// width-5 NAF of a canonical scalar, 256 digits (dalek Scalar::non_adjacent_form(5))
static inline void sc_naf5(int8_t naf[257], const sc &s) {
    memset(naf, 0, 257);
    uint64_t x[5] = {s.v[0], s.v[1], s.v[2], s.v[3], 0};
    int pos = 0;
    uint64_t carry = 0;
    while (pos < 256) {
        int idx = pos >> 6, off = pos & 63;
        uint64_t bits = (off < 59) ? (x[idx] >> off) : ((x[idx] >> off) | (x[idx + 1] << (64 - off)));
        uint64_t window = carry + (bits & 31);
        if ((window & 1) == 0) { pos += 1; continue; }
        if (window < 16) { carry = 0; naf[pos] = (int8_t)window; }
        else { carry = 1; naf[pos] = (int8_t)((int)window - 32); }
        pos += 5;
    }
    if (carry) naf[256] = 1;
}

// a*P + b*Q with interleaved width-5 NAFs (what dalek's Straus does for tiny vartime MSMs)
static inline ge ge_double_scalarmul_vartime(const sc &a, const ge &P, const sc &b, const ge &Q) {
    int8_t na[257], nb[257];
    sc_naf5(na, a); sc_naf5(nb, b);
    ge tp[8], tq[8];
    ge P2 = ge_dbl(P), Q2 = ge_dbl(Q);
    tp[0] = P; tq[0] = Q;
    for (int i = 1; i < 8; i++) { tp[i] = ge_add(tp[i - 1], P2); tq[i] = ge_add(tq[i - 1], Q2); }
    int i = 256;
    while (i >= 0 && na[i] == 0 && nb[i] == 0) i--;
    ge r = ge_identity();
    for (; i >= 0; i--) {
        r = ge_dbl(r);
        if (na[i] > 0) r = ge_add(r, tp[na[i] >> 1]);
        else if (na[i] < 0) r = ge_sub(r, tp[(-na[i]) >> 1]);
        if (nb[i] > 0) r = ge_add(r, tq[nb[i] >> 1]);
        else if (nb[i] < 0) r = ge_sub(r, tq[(-nb[i]) >> 1]);
    }
    return r;
}

static inline bytes32 ge_compress32(const ge &p) {
    bytes32 o;
    ge_compress(o.data(), p);
    return o;
}

// InnerProductProof::create. G, H, a, b are consumed (folded in place). n must be a power of two.
static inline ipp_proof ipp_create(transcript &tr, const ge &Q, const std::vector<sc> &Gf, const std::vector<sc> &Hf,
                                   std::vector<ge> G, std::vector<ge> H, std::vector<sc> a, std::vector<sc> b) {
    size_t n = G.size();
    tr.innerproduct_domain_sep(n);
    ipp_proof pf;
    bool first = true;
    while (n != 1) {
        n /= 2;
        sc c_L = sc_inner_product(&a[0], &b[n], n);
        sc c_R = sc_inner_product(&a[n], &b[0], n);
        std::vector<sc> sl(2 * n + 1), sr(2 * n + 1);
        std::vector<ge> pl(2 * n + 1), pr(2 * n + 1);
        for (size_t i = 0; i < n; i++) {
            sl[i] = first ? sc_mul(a[i], Gf[n + i]) : a[i];          pl[i] = G[n + i];
            sl[n + i] = first ? sc_mul(b[n + i], Hf[i]) : b[n + i];  pl[n + i] = H[i];
            sr[i] = first ? sc_mul(a[n + i], Gf[i]) : a[n + i];      pr[i] = G[i];
            sr[n + i] = first ? sc_mul(b[i], Hf[n + i]) : b[i];      pr[n + i] = H[n + i];
        }
        sl[2 * n] = c_L; pl[2 * n] = Q;
        sr[2 * n] = c_R; pr[2 * n] = Q;
        bytes32 Lc = ge_compress32(msm_pippenger_serial(sl.data(), pl.data(), 2 * n + 1));
        bytes32 Rc = ge_compress32(msm_pippenger_serial(sr.data(), pr.data(), 2 * n + 1));
        pf.L.push_back(Lc); pf.R.push_back(Rc);
        tr.append_point("L", Lc.data());
        tr.append_point("R", Rc.data());
        sc u = tr.challenge_scalar("u");
        sc u_inv = sc_invert(u);
        for (size_t i = 0; i < n; i++) {
            a[i] = sc_add(sc_mul(a[i], u), sc_mul(u_inv, a[n + i]));
            b[i] = sc_add(sc_mul(b[i], u_inv), sc_mul(u, b[n + i]));
            sc g0 = u_inv, g1 = u, h0 = u, h1 = u_inv;
            if (first) {
                g0 = sc_mul(u_inv, Gf[i]); g1 = sc_mul(u, Gf[n + i]);
                h0 = sc_mul(u, Hf[i]);     h1 = sc_mul(u_inv, Hf[n + i]);
            }
            G[i] = ge_double_scalarmul_vartime(g0, G[i], g1, G[n + i]);
            H[i] = ge_double_scalarmul_vartime(h0, H[i], h1, H[n + i]);
        }
        a.resize(n); b.resize(n); G.resize(n); H.resize(n);
        first = false;
    }
    pf.a = a[0]; pf.b = b[0];
    return pf;
}

static inline std::vector<uint8_t> ipp_to_bytes(const ipp_proof &pf) {
    std::vector<uint8_t> out;
    for (size_t i = 0; i < pf.L.size(); i++) {
        out.insert(out.end(), pf.L[i].begin(), pf.L[i].end());
        out.insert(out.end(), pf.R[i].begin(), pf.R[i].end());
    }
    uint8_t t[32];
    sc_tobytes(t, pf.a); out.insert(out.end(), t, t + 32);
    sc_tobytes(t, pf.b); out.insert(out.end(), t, t + 32);
    return out;
}

// InnerProductProof::from_bytes; false = FormatError
static inline bool ipp_from_bytes(ipp_proof &pf, const uint8_t *p, size_t len) {
    if (len % 32 != 0) return false;
    size_t ne = len / 32;
    if (ne < 2) return false;
    if ((ne - 2) % 2 != 0) return false;
    size_t lg_n = (ne - 2) / 2;
    if (lg_n >= 32) return false;
    pf.L.resize(lg_n); pf.R.resize(lg_n);
    for (size_t i = 0; i < lg_n; i++) {
        memcpy(pf.L[i].data(), p + 64 * i, 32);
        memcpy(pf.R[i].data(), p + 64 * i + 32, 32);
    }
    if (!sc_from_canonical(pf.a, p + 64 * lg_n)) return false;
    if (!sc_from_canonical(pf.b, p + 64 * lg_n + 32)) return false;
    return true;
}

// InnerProductProof::verification_scalars; false = VerificationError
static inline bool ipp_verification_scalars(const ipp_proof &pf, size_t n, transcript &tr, std::vector<sc> &u_sq,
                                            std::vector<sc> &u_inv_sq, std::vector<sc> &s) {
    size_t lg_n = pf.L.size();
    if (lg_n >= 32) return false;
    if (n != ((size_t)1 << lg_n)) return false;
    tr.innerproduct_domain_sep(n);
    std::vector<sc> ch(lg_n);
    for (size_t i = 0; i < lg_n; i++) {
        if (!tr.validate_and_append_point("L", pf.L[i].data())) return false;
        if (!tr.validate_and_append_point("R", pf.R[i].data())) return false;
        ch[i] = tr.challenge_scalar("u");
    }
    std::vector<sc> ch_inv = ch;
    sc allinv = sc_batch_invert(ch_inv);
    u_sq.resize(lg_n); u_inv_sq.resize(lg_n);
    for (size_t i = 0; i < lg_n; i++) { u_sq[i] = sc_mul(ch[i], ch[i]); u_inv_sq[i] = sc_mul(ch_inv[i], ch_inv[i]); }
    s.resize(n);
    s[0] = allinv;
    for (size_t i = 1; i < n; i++) {
        size_t lg_i = 63 - __builtin_clzll((unsigned long long)i);
        size_t k = (size_t)1 << lg_i;
        s[i] = sc_mul(s[i - k], u_sq[(lg_n - 1) - lg_i]);
    }
    return true;
}

}  // namespace orc
