// ORACLE — test infrastructure only (see fe.h header). C ABI over the CPU restatement so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can drive it
// through ctypes. Nothing under dusk-blindbidproof_b200/ links, loads or calls this library.
//
// Conventions: scalars are 32-byte little-endian, points are 32-byte compressed Ristretto unless a
// function says "ext" (128 bytes = X,Y,Z,T as 4 canonical field encodings).
#include <map>
#include <thread>
#include <string>
#include "blindbid.h"
#include "rangeproof.h"
#include <chrono>

using namespace orc;

static inline sc sc_load_reduced(const uint8_t *p) { return sc_from_bytes_mod_order(p); }

static inline void ge_to_ext(uint8_t out[128], const ge &p) {
    fe_tobytes(out, p.X); fe_tobytes(out + 32, p.Y); fe_tobytes(out + 64, p.Z); fe_tobytes(out + 96, p.T);
}
static inline ge ge_from_ext(const uint8_t in[128]) {
    return ge{fe_frombytes(in), fe_frombytes(in + 32), fe_frombytes(in + 64), fe_frombytes(in + 96)};
}

extern "C" {

// ---------------- field ----------------
void orc_fe_mul(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { fe_tobytes(out, fe_mul(fe_frombytes(a), fe_frombytes(b))); }
void orc_fe_add(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { fe_tobytes(out, fe_add(fe_frombytes(a), fe_frombytes(b))); }
void orc_fe_sub(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { fe_tobytes(out, fe_sub(fe_frombytes(a), fe_frombytes(b))); }
void orc_fe_invert(uint8_t out[32], const uint8_t a[32]) { fe_tobytes(out, fe_invert(fe_frombytes(a))); }
// constants in the order d, 2d, sqrt_m1, sqrt_ad_minus_one, invsqrt_a_minus_d, one_minus_d_sq, d_minus_one_sq
void orc_fe_constants(uint8_t out[7 * 32]) {
    const fe_consts &K = fe_constants();
    fe_tobytes(out, K.d); fe_tobytes(out + 32, K.d2); fe_tobytes(out + 64, K.sqrt_m1); fe_tobytes(out + 96, K.sqrt_ad_minus_one);
    fe_tobytes(out + 128, K.invsqrt_a_minus_d); fe_tobytes(out + 160, K.one_minus_d_sq); fe_tobytes(out + 192, K.d_minus_one_sq);
}
int orc_fe_sqrt_ratio_i(uint8_t out[32], const uint8_t u[32], const uint8_t v[32]) {
    fe r;
    bool ok = fe_sqrt_ratio_i(r, fe_frombytes(u), fe_frombytes(v), fe_constants().sqrt_m1);
    fe_tobytes(out, r);
    return ok ? 1 : 0;
}

// ---------------- scalars ----------------
void orc_sc_mul(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { sc_tobytes(out, sc_mul(sc_load_reduced(a), sc_load_reduced(b))); }
void orc_sc_add(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { sc_tobytes(out, sc_add(sc_load_reduced(a), sc_load_reduced(b))); }
void orc_sc_sub(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { sc_tobytes(out, sc_sub(sc_load_reduced(a), sc_load_reduced(b))); }
void orc_sc_invert(uint8_t out[32], const uint8_t a[32]) { sc_tobytes(out, sc_invert(sc_load_reduced(a))); }
void orc_sc_from_wide(uint8_t out[32], const uint8_t in[64]) { sc_tobytes(out, sc_from_wide(in)); }
void orc_sc_reduce32(uint8_t out[32], const uint8_t in[32]) { sc_tobytes(out, sc_from_bytes_mod_order(in)); }
int orc_sc_is_canonical(const uint8_t in[32]) { sc t; return sc_from_canonical(t, in) ? 1 : 0; }
void orc_sc_batch_invert(uint8_t *inout, size_t n, uint8_t allinv[32]) {
    std::vector<sc> xs(n);
    for (size_t i = 0; i < n; i++) xs[i] = sc_load_reduced(inout + 32 * i);
    sc r = sc_batch_invert(xs);
    for (size_t i = 0; i < n; i++) sc_tobytes(inout + 32 * i, xs[i]);
    sc_tobytes(allinv, r);
}

// ---------------- hashes / transcript ----------------
void orc_sha512(uint8_t out[64], const uint8_t *in, size_t n) { sha512(out, in, n); }
void orc_sha3_512(uint8_t out[64], const uint8_t *in, size_t n) { sha3_512(out, in, n); }
void orc_shake256(uint8_t *out, size_t outlen, const uint8_t *in, size_t n) {
    shake256 s;
    s.absorb(in, n);
    s.squeeze(out, outlen);
}
// Merlin: Transcript::new(label); a list of (label,msg) appends; challenge_bytes(clabel, outlen)
void orc_merlin_simple(uint8_t *out, size_t outlen, const char *label, const char *alabel, const uint8_t *msg, size_t msglen, const char *clabel) {
    transcript t(label);
    t.append_message(alabel, msg, msglen);
    t.challenge_bytes(clabel, out, outlen);
}
// TranscriptRng: new(label) -> build_rng -> rekey(wlabel, witness) -> finalize(ext32) -> fill(outlen)
void orc_merlin_rng(uint8_t *out, size_t outlen, const char *label, const char *wlabel, const uint8_t *w, size_t wlen, const uint8_t ext32[32]) {
    transcript t(label);
    transcript_rng_builder b = t.build_rng();
    b.rekey_with_witness_bytes(wlabel, w, wlen);
    transcript_rng r = b.finalize(ext32);
    r.fill_bytes(out, outlen);
}

// ---------------- group ----------------
int orc_ge_decompress_ext(uint8_t out[128], const uint8_t in[32]) {
    ge p;
    if (!ge_decompress(p, in)) return 0;
    ge_to_ext(out, p);
    return 1;
}
void orc_ge_compress_ext(uint8_t out[32], const uint8_t in[128]) { ge_compress(out, ge_from_ext(in)); }
// decompress + recompress (round trip); 0 if the encoding is invalid
int orc_ge_roundtrip(uint8_t out[32], const uint8_t in[32]) {
    ge p;
    if (!ge_decompress(p, in)) return 0;
    ge_compress(out, p);
    return 1;
}
void orc_ge_from_uniform(uint8_t out[32], const uint8_t in[64]) { ge_compress(out, ge_from_uniform_bytes(in)); }
void orc_ge_from_uniform_ext(uint8_t out[128], const uint8_t in[64]) { ge_to_ext(out, ge_from_uniform_bytes(in)); }
void orc_ge_basepoint(uint8_t out[32]) { ge_compress(out, ge_basepoint()); }
int orc_ge_scalarmul(uint8_t out[32], const uint8_t s[32], const uint8_t p[32]) {
    ge P;
    if (!ge_decompress(P, p)) return 0;
    ge_compress(out, ge_scalarmul(sc_load_reduced(s), P));
    return 1;
}
int orc_ge_add(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) {
    ge A, B;
    if (!ge_decompress(A, a) || !ge_decompress(B, b)) return 0;
    ge_compress(out, ge_add(A, B));
    return 1;
}
int orc_ge_double(uint8_t out[32], const uint8_t a[32]) {
    ge A;
    if (!ge_decompress(A, a)) return 0;
    ge_compress(out, ge_dbl(A));
    return 1;
}
// algo: 0 naive, 1 pippenger. Returns 0 if any point fails to decompress (optional_multiscalar_mul -> None)
int orc_msm(uint8_t out[32], const uint8_t *scalars, const uint8_t *points, size_t n, int algo, int threads) {
    std::vector<sc> s(n);
    std::vector<ge> p(n);
    for (size_t i = 0; i < n; i++) {
        s[i] = sc_load_reduced(scalars + 32 * i);
        if (!ge_decompress(p[i], points + 32 * i)) return 0;
    }
    ge r = algo == 0 ? msm_naive(s.data(), p.data(), n) : msm_pippenger(s.data(), p.data(), n, threads);
    ge_compress(out, r);
    return 1;
}
// MSM over extended-coordinate inputs (n x 128 B), timed separately from decompression. Returns seconds.
double orc_msm_ext(uint8_t out[32], const uint8_t *scalars, const uint8_t *points_ext, size_t n, int threads) {
    std::vector<sc> s(n);
    std::vector<ge> p(n);
    for (size_t i = 0; i < n; i++) { s[i] = sc_load_reduced(scalars + 32 * i); p[i] = ge_from_ext(points_ext + 128 * i); }
    auto t0 = std::chrono::steady_clock::now();
    ge r = msm_pippenger(s.data(), p.data(), n, threads);
    auto t1 = std::chrono::steady_clock::now();
    ge_compress(out, r);
    return std::chrono::duration<double>(t1 - t0).count();
}

// n uniform points from n 64-byte blocks (from_uniform_bytes), extended coordinates, on `threads` host threads: builds the
// CPU arm's inputs of bench.py (the same SHAKE256 stream the GPU arm maps to the group with its own codec)
void orc_points_from_uniform_ext_mt(uint8_t *out, const uint8_t *in, size_t n, int threads) {
    if (threads < 1) threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([=] {
            for (size_t i = n * t / threads; i < n * (t + 1) / threads; i++) ge_to_ext(out + 128 * i, ge_from_uniform_bytes(in + 64 * i));
        });
    for (auto &x : th) x.join();
}

// ---------------- generators / constants ----------------
void orc_pedersen_gens(uint8_t B[32], uint8_t B_blinding[32]) {
    pedersen_gens pc;
    ge_compress(B, pc.B);
    ge_compress(B_blinding, pc.B_blinding);
}
// which = 'G' or 'H'; writes count compressed points of the party's chain
void orc_bp_gens(uint8_t *out, int which, uint32_t party, size_t count) {
    std::vector<ge> g;
    generators_chain(g, (char)which, party, count);
    for (size_t i = 0; i < count; i++) ge_compress(out + 32 * i, g[i]);
}
void orc_bp_gens_ext(uint8_t *out, int which, uint32_t party, size_t count) {
    std::vector<ge> g;
    generators_chain(g, (char)which, party, count);
    for (size_t i = 0; i < count; i++) ge_to_ext(out + 128 * i, g[i]);
}
void orc_mimc_constants(uint8_t out[90 * 32]) {
    const std::vector<sc> &c = mimc_constants();
    for (size_t i = 0; i < MIMC_ROUNDS; i++) sc_tobytes(out + 32 * i, c[i]);
}
void orc_mimc_hash(uint8_t out[32], const uint8_t left[32], const uint8_t right[32]) {
    sc_tobytes(out, mimc_hash(sc_load_reduced(left), sc_load_reduced(right)));
}

// ---------------- blind bid ----------------
// circuit shape for a public list of length L: multipliers n1, constraints q, commitments m
void orc_blindbid_shape(size_t L, size_t out[3]) {
    transcript tr("BlindBidProofGadget");
    verifier ve(tr);
    bytes32 z;
    z.fill(0);
    std::vector<variable> vars, tv;
    for (int i = 0; i < 4; i++) vars.push_back(ve.commit(z));
    for (size_t i = 0; i < L; i++) tv.push_back(ve.commit(z));
    std::vector<lincomb> lv(L, lincomb(sc_one()));
    proof_gadget(ve, lincomb(vars[0]), lincomb(vars[1]), lincomb(vars[3]), lincomb(sc_one()), lincomb(sc_one()), lincomb(sc_one()),
                 mimc_constants(), tv, lv);
    out[0] = ve.num_vars; out[1] = ve.constraints.size(); out[2] = ve.V.size();
}

// Proof::prove. blindings = (4+L) x 32. Outputs: proof bytes (R1CSProof::to_bytes, `versioned` layout switch),
// commitments 4 x 32, t_c L x 32. Returns r1cs_error; *proof_len in = capacity, out = length.
int orc_blindbid_prove(const uint8_t d[32], const uint8_t k[32], const uint8_t y[32], const uint8_t y_inv[32], const uint8_t q[32],
                       const uint8_t z_img[32], const uint8_t seed[32], const uint8_t *pub_list, size_t L, uint64_t toggle,
                       const uint8_t *blindings, const uint8_t rng32[32], int versioned, uint8_t *proof_out, size_t *proof_len,
                       uint8_t *commitments_out, uint8_t *t_c_out) {
    if (L == 0) return R1CS_FORMAT_ERROR;
    std::vector<sc> pl(L), bl(4 + L);
    for (size_t i = 0; i < L; i++) pl[i] = sc_from_bits(pub_list + 32 * i);
    for (size_t i = 0; i < 4 + L; i++) bl[i] = sc_load_reduced(blindings + 32 * i);
    blindbid_proof out;
    int rc = blindbid_prove(sc_load_reduced(d), sc_load_reduced(k), sc_load_reduced(y), sc_load_reduced(y_inv), sc_load_reduced(q),
                            sc_load_reduced(z_img), sc_load_reduced(seed), pl, toggle, bl, rng32, out);
    if (rc != R1CS_OK) return rc;
    std::vector<uint8_t> pb = r1cs_proof_to_bytes(out.proof, versioned != 0);
    if (pb.size() > *proof_len) return R1CS_FORMAT_ERROR;
    memcpy(proof_out, pb.data(), pb.size());
    *proof_len = pb.size();
    for (int i = 0; i < 4; i++) memcpy(commitments_out + 32 * i, out.commitments[i].data(), 32);
    for (size_t i = 0; i < L; i++) memcpy(t_c_out + 32 * i, out.t_c[i].data(), 32);
    return R1CS_OK;
}

// Verify::verify. Returns 0 = accept, negative r1cs_error otherwise. mega_scalars_out (optional) receives the
// assembled mega-check scalars (count returned through *n_mega).
int orc_blindbid_verify(const uint8_t *proof, size_t proof_len, int versioned, const uint8_t *commitments, size_t nc, const uint8_t *t_c,
                        size_t nt, const uint8_t score[32], const uint8_t z_img[32], const uint8_t seed[32], const uint8_t *pub_list,
                        size_t L, const uint8_t rng32[32], int threads, uint8_t *mega_scalars_out, size_t *n_mega) {
    blindbid_proof p;
    int rc = r1cs_proof_from_bytes(p.proof, proof, proof_len, versioned != 0);
    if (rc != R1CS_OK) return rc;
    p.commitments.resize(nc);
    p.t_c.resize(nt);
    for (size_t i = 0; i < nc; i++) memcpy(p.commitments[i].data(), commitments + 32 * i, 32);
    for (size_t i = 0; i < nt; i++) memcpy(p.t_c[i].data(), t_c + 32 * i, 32);
    std::vector<sc> pl(L);
    for (size_t i = 0; i < L; i++) pl[i] = sc_from_bits(pub_list + 32 * i);
    std::vector<sc> mega;
    rc = blindbid_verify(p, sc_load_reduced(score), sc_load_reduced(z_img), sc_load_reduced(seed), pl, rng32, threads,
                         mega_scalars_out ? &mega : nullptr);
    if (mega_scalars_out && n_mega) {
        size_t cap = *n_mega;
        *n_mega = mega.size();
        for (size_t i = 0; i < mega.size() && i < cap; i++) sc_tobytes(mega_scalars_out + 32 * i, mega[i]);
    }
    return rc;
}

// ---------------- generic R1CS (flattened circuits) and the standalone inner-product argument ----------------
// The circuit as ConstraintSystem::{multiply, constrain} leave it (src/gadgets.rs:30,53): constraint j owns terms
// con_ptr[j] .. con_ptr[j+1], term = (kind << 28 | index, coefficient). Feeds the oracle's generic prover / verifier
// (r1cs.h) directly, so that the product's bbp_r1cs_prove / bbp_r1cs_verify can be compared on arbitrary circuits.
static bool load_flat(std::vector<lincomb> &cons, size_t q, const uint32_t *con_ptr, const uint32_t *term_var, const uint8_t *term_coeff) {
    cons.resize(q);
    for (size_t j = 0; j < q; j++)
        for (uint32_t t = con_ptr[j]; t < con_ptr[j + 1]; t++) {
            uint32_t kind = term_var[t] >> 28, idx = term_var[t] & 0x0fffffffu;
            if (kind > 4) return false;
            sc c;
            if (!sc_from_canonical(c, term_coeff + 32 * (size_t)t)) return false;
            cons[j].terms.push_back({variable{(var_kind)kind, idx}, c});
        }
    return true;
}
static const bulletproof_gens &flat_gens(size_t cap) {
    static std::map<size_t, bulletproof_gens *> cache;
    auto it = cache.find(cap);
    if (it == cache.end()) it = cache.emplace(cap, new bulletproof_gens(cap, 1)).first;
    return *it->second;
}
int orc_r1cs_prove_flat(const uint8_t *label, size_t label_len, size_t gens_capacity, size_t n_mul, size_t m, size_t q, const uint32_t *con_ptr,
                        const uint32_t *term_var, const uint8_t *term_coeff, const uint8_t *a_L, const uint8_t *a_R, const uint8_t *a_O,
                        const uint8_t *v, const uint8_t *v_blinding, const uint8_t rng32[32], int versioned, uint8_t *V_out, uint8_t *proof_out,
                        size_t *proof_len, uint8_t challenge_after[32]) {
    std::string lbl((const char *)label, label_len);
    transcript tr(lbl.c_str());
    pedersen_gens pc;
    prover P(pc, tr);
    for (size_t i = 0; i < m; i++) {
        bytes32 V;
        P.commit(sc_load_reduced(v + 32 * i), sc_load_reduced(v_blinding + 32 * i), V);
        memcpy(V_out + 32 * i, V.data(), 32);
    }
    for (size_t i = 0; i < n_mul; i++) {
        P.a_L.push_back(sc_load_reduced(a_L + 32 * i)); P.a_R.push_back(sc_load_reduced(a_R + 32 * i)); P.a_O.push_back(sc_load_reduced(a_O + 32 * i));
    }
    if (!load_flat(P.constraints, q, con_ptr, term_var, term_coeff)) return R1CS_FORMAT_ERROR;
    r1cs_proof out;
    int rc = P.prove(flat_gens(gens_capacity), rng32, out);
    if (rc != R1CS_OK) return rc;
    std::vector<uint8_t> pb = r1cs_proof_to_bytes(out, versioned != 0);
    if (pb.size() > *proof_len) return R1CS_FORMAT_ERROR;
    memcpy(proof_out, pb.data(), pb.size());
    *proof_len = pb.size();
    if (challenge_after) tr.challenge_bytes("after", challenge_after, 32);   // pins the transcript state the proof leaves behind
    return R1CS_OK;
}
int orc_r1cs_verify_flat(const uint8_t *label, size_t label_len, size_t gens_capacity, size_t n_mul, size_t m, size_t q, const uint32_t *con_ptr,
                         const uint32_t *term_var, const uint8_t *term_coeff, const uint8_t *proof, size_t proof_len, int versioned,
                         const uint8_t *V, const uint8_t rng32[32], uint8_t challenge_after[32]) {
    std::string lbl((const char *)label, label_len);
    transcript tr(lbl.c_str());
    verifier ve(tr);
    for (size_t i = 0; i < m; i++) {
        bytes32 c;
        memcpy(c.data(), V + 32 * i, 32);
        ve.commit(c);
    }
    ve.num_vars = n_mul;
    if (!load_flat(ve.constraints, q, con_ptr, term_var, term_coeff)) return R1CS_FORMAT_ERROR;
    r1cs_proof p;
    int rc = r1cs_proof_from_bytes(p, proof, proof_len, versioned != 0);
    if (rc != R1CS_OK) return rc;
    pedersen_gens pc;
    rc = ve.verify(p, pc, flat_gens(gens_capacity), rng32);
    if (rc == R1CS_OK && challenge_after) tr.challenge_bytes("after", challenge_after, 32);
    return rc;
}
// InnerProductProof::create with Q = w * B over the first n generators of BulletproofGens::new(n, 1)
int orc_ipp_create(const uint8_t *label, size_t label_len, const uint8_t w[32], const uint8_t *Gf, const uint8_t *Hf, const uint8_t *a,
                   const uint8_t *b, size_t n, uint8_t *out, uint8_t challenge_after[32]) {
    std::string lbl((const char *)label, label_len);
    transcript tr(lbl.c_str());
    pedersen_gens pc;
    const bulletproof_gens &bp = flat_gens(n);
    std::vector<sc> gf(n), hf(n), av(n), bv(n);
    for (size_t i = 0; i < n; i++) {
        gf[i] = sc_load_reduced(Gf + 32 * i); hf[i] = sc_load_reduced(Hf + 32 * i); av[i] = sc_load_reduced(a + 32 * i); bv[i] = sc_load_reduced(b + 32 * i);
    }
    std::vector<ge> G(bp.G[0].begin(), bp.G[0].begin() + n), H(bp.H[0].begin(), bp.H[0].begin() + n);
    ge Q = ge_scalarmul(sc_load_reduced(w), pc.B);
    ipp_proof pf = ipp_create(tr, Q, gf, hf, std::move(G), std::move(H), std::move(av), std::move(bv));
    std::vector<uint8_t> bytes = ipp_to_bytes(pf);
    memcpy(out, bytes.data(), bytes.size());
    if (challenge_after) tr.challenge_bytes("after", challenge_after, 32);
    return 0;
}

// ---------------- aggregated range proof (config 5) ----------------
int orc_rangeproof_prove(const uint64_t *values, const uint8_t *blindings, size_t m, size_t nbits, const uint8_t *rng_stream_seed32,
                         uint8_t *proof_out, size_t *proof_len, uint8_t *commitments_out) {
    std::vector<sc> bl(m);
    for (size_t i = 0; i < m; i++) bl[i] = sc_load_reduced(blindings + 32 * i);
    std::vector<uint64_t> v(values, values + m);
    std::vector<uint8_t> pb;
    std::vector<bytes32> V;
    int rc = rangeproof_prove_multiple(v, bl, nbits, rng_stream_seed32, pb, V);
    if (rc != 0) return rc;
    if (pb.size() > *proof_len) return -2;
    memcpy(proof_out, pb.data(), pb.size());
    *proof_len = pb.size();
    for (size_t i = 0; i < m; i++) memcpy(commitments_out + 32 * i, V[i].data(), 32);
    return 0;
}
int orc_rangeproof_verify(const uint8_t *proof, size_t proof_len, const uint8_t *commitments, size_t m, size_t nbits, const uint8_t rng32[32],
                          int threads) {
    std::vector<bytes32> V(m);
    for (size_t i = 0; i < m; i++) memcpy(V[i].data(), commitments + 32 * i, 32);
    return rangeproof_verify_multiple(proof, proof_len, V, nbits, rng32, threads);
}

}  // extern "C"
