// Host-side Keccak-f[1600] sponge functions and the STROBE-128 / Merlin transcript layer of the product.
// Stands in for the crates the reference links: keccak 0.1.0 / sha3 0.8.2 (generators: bulletproofs GeneratorsChain,
// PedersenGens), merlin 1.3.0 (Transcript, TranscriptRng) — Cargo.lock:366-367,648-649,399-400 — and sha2 0.8.0's
// SHA-512 used for the MiMC constants at src/blindbid/mod.rs:11,18. Transcripts are sequential and tiny (SURVEY.md
// §2.2 U9), so they stay on the host, one per proof, pipelined ahead of the GPU work.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "sc25519.cuh"

namespace bbp {

// Keccak-f[1600], lane-wise and unrolled per round (the transcript RNG draws ~3 k permutations per proof, so this is the
// host-side hot loop of the prover)
inline uint64_t keccak_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
inline void keccak_f1600(uint64_t s[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    uint64_t a00 = s[0], a01 = s[1], a02 = s[2], a03 = s[3], a04 = s[4], a05 = s[5], a06 = s[6], a07 = s[7], a08 = s[8], a09 = s[9], a10 = s[10],
             a11 = s[11], a12 = s[12], a13 = s[13], a14 = s[14], a15 = s[15], a16 = s[16], a17 = s[17], a18 = s[18], a19 = s[19], a20 = s[20],
             a21 = s[21], a22 = s[22], a23 = s[23], a24 = s[24];
    for (int r = 0; r < 24; r++) {
        // theta
        uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                 c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
        uint64_t d0 = c4 ^ keccak_rotl(c1, 1), d1 = c0 ^ keccak_rotl(c2, 1), d2 = c1 ^ keccak_rotl(c3, 1), d3 = c2 ^ keccak_rotl(c4, 1),
                 d4 = c3 ^ keccak_rotl(c0, 1);
        a00 ^= d0; a05 ^= d0; a10 ^= d0; a15 ^= d0; a20 ^= d0;
        a01 ^= d1; a06 ^= d1; a11 ^= d1; a16 ^= d1; a21 ^= d1;
        a02 ^= d2; a07 ^= d2; a12 ^= d2; a17 ^= d2; a22 ^= d2;
        a03 ^= d3; a08 ^= d3; a13 ^= d3; a18 ^= d3; a23 ^= d3;
        a04 ^= d4; a09 ^= d4; a14 ^= d4; a19 ^= d4; a24 ^= d4;
        // rho + pi: b[y][2x+3y] = rotl(a[x][y]) with lanes indexed a[x + 5y]
        uint64_t b00 = a00, b10 = keccak_rotl(a01, 1), b20 = keccak_rotl(a02, 62), b05 = keccak_rotl(a03, 28), b15 = keccak_rotl(a04, 27),
                 b16 = keccak_rotl(a05, 36), b01 = keccak_rotl(a06, 44), b11 = keccak_rotl(a07, 6), b21 = keccak_rotl(a08, 55), b06 = keccak_rotl(a09, 20),
                 b07 = keccak_rotl(a10, 3), b17 = keccak_rotl(a11, 10), b02 = keccak_rotl(a12, 43), b12 = keccak_rotl(a13, 25), b22 = keccak_rotl(a14, 39),
                 b23 = keccak_rotl(a15, 41), b08 = keccak_rotl(a16, 45), b18 = keccak_rotl(a17, 15), b03 = keccak_rotl(a18, 21), b13 = keccak_rotl(a19, 8),
                 b14 = keccak_rotl(a20, 18), b24 = keccak_rotl(a21, 2), b09 = keccak_rotl(a22, 61), b19 = keccak_rotl(a23, 56), b04 = keccak_rotl(a24, 14);
        // chi
        a00 = b00 ^ (~b01 & b02); a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
        a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
        a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
        a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
        a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
        a00 ^= RC[r];
    }
    s[0] = a00; s[1] = a01; s[2] = a02; s[3] = a03; s[4] = a04; s[5] = a05; s[6] = a06; s[7] = a07; s[8] = a08; s[9] = a09; s[10] = a10; s[11] = a11;
    s[12] = a12; s[13] = a13; s[14] = a14; s[15] = a15; s[16] = a16; s[17] = a17; s[18] = a18; s[19] = a19; s[20] = a20; s[21] = a21; s[22] = a22;
    s[23] = a23; s[24] = a24;
}

struct keccak_sponge {
    uint64_t st[25];
    size_t rate, pos;
    uint8_t pad;
    bool squeezing;
    keccak_sponge(size_t rate_bytes, uint8_t pad_byte) : rate(rate_bytes), pos(0), pad(pad_byte), squeezing(false) { memset(st, 0, sizeof st); }
    void absorb(const void *data, size_t n) {
        const uint8_t *d = (const uint8_t *)data;
        uint8_t *s = (uint8_t *)st;
        while (n--) {
            s[pos++] ^= *d++;
            if (pos == rate) { keccak_f1600(st); pos = 0; }
        }
    }
    void squeeze(void *out, size_t n) {
        uint8_t *o = (uint8_t *)out;
        uint8_t *s = (uint8_t *)st;
        if (!squeezing) {
            s[pos] ^= pad;
            s[rate - 1] ^= 0x80;
            keccak_f1600(st);
            pos = 0;
            squeezing = true;
        }
        while (n--) {
            if (pos == rate) { keccak_f1600(st); pos = 0; }
            *o++ = s[pos++];
        }
    }
};
inline keccak_sponge shake256_new() { return keccak_sponge(136, 0x1f); }
inline void sha3_512(uint8_t out[64], const void *in, size_t n) {
    keccak_sponge s(72, 0x06);
    s.absorb(in, n);
    s.squeeze(out, 64);
}

// FIPS 180-4 SHA-512 (one-shot), constants generated from their definition at start-up
struct sha512_tables {
    uint64_t K[80], H0[8];
    sha512_tables() {
        // K = frac(cbrt(prime_i)) * 2^64, H0 = frac(sqrt(prime_i)) * 2^64, by integer Newton iterations on 192/128-bit values
        int primes[80], np = 0;
        for (int c = 2; np < 80; c++) {
            bool ok = true;
            for (int i = 0; i < np; i++) if (c % primes[i] == 0) { ok = false; break; }
            if (ok) primes[np++] = c;
        }
        for (int i = 0; i < 80; i++) {
            // floor(cbrt(p * 2^192)) mod 2^64 via binary search on a 72-bit root held in unsigned __int128
            unsigned __int128 lo = 0, hi = (unsigned __int128)1 << 72;
            while (lo + 1 < hi) {
                unsigned __int128 mid = (lo + hi) >> 1;
                if (cube_leq(mid, (uint64_t)primes[i])) lo = mid; else hi = mid;
            }
            K[i] = (uint64_t)lo;
        }
        for (int i = 0; i < 8; i++) {
            unsigned __int128 lo = 0, hi = (unsigned __int128)1 << 68;
            while (lo + 1 < hi) {
                unsigned __int128 mid = (lo + hi) >> 1;
                if (square_leq(mid, (uint64_t)primes[i])) lo = mid; else hi = mid;
            }
            H0[i] = (uint64_t)lo;
        }
    }
    // mid^3 <= p * 2^192 ?  (mid < 2^72): compare through 256-bit arithmetic in 64-bit limbs
    static void mul_limbs(const uint64_t *a, int na, const uint64_t *b, int nb, uint64_t *r) {
        for (int i = 0; i < na + nb; i++) r[i] = 0;
        for (int i = 0; i < na; i++) {
            unsigned __int128 c = 0;
            for (int j = 0; j < nb; j++) {
                c += (unsigned __int128)a[i] * b[j] + r[i + j];
                r[i + j] = (uint64_t)c;
                c >>= 64;
            }
            r[i + nb] = (uint64_t)c;
        }
    }
    static bool cube_leq(unsigned __int128 m, uint64_t p) {
        uint64_t a[2] = {(uint64_t)m, (uint64_t)(m >> 64)}, sq[4], cu[6];
        mul_limbs(a, 2, a, 2, sq);
        mul_limbs(sq, 4, a, 2, cu);
        // target = p << 192 : limb 3 = p, others zero
        if (cu[5] || cu[4]) return false;
        if (cu[3] != p) return cu[3] < p;
        return (cu[2] | cu[1] | cu[0]) == 0;
    }
    static bool square_leq(unsigned __int128 m, uint64_t p) {
        uint64_t a[2] = {(uint64_t)m, (uint64_t)(m >> 64)}, sq[4];
        mul_limbs(a, 2, a, 2, sq);
        // target = p << 128
        if (sq[3]) return false;
        if (sq[2] != p) return sq[2] < p;
        return (sq[1] | sq[0]) == 0;
    }
};
inline void sha512(uint8_t out[64], const void *in, size_t n) {
    static const sha512_tables T;
    auto rotr = [](uint64_t x, int k) { return (x >> k) | (x << (64 - k)); };
    uint64_t h[8];
    memcpy(h, T.H0, sizeof h);
    size_t total = ((n + 17 + 127) / 128) * 128;
    std::vector<uint8_t> m(total, 0);
    memcpy(m.data(), in, n);
    m[n] = 0x80;
    uint64_t bits = (uint64_t)n * 8;
    for (int i = 0; i < 8; i++) m[total - 1 - i] = (uint8_t)(bits >> (8 * i));
    for (size_t off = 0; off < total; off += 128) {
        uint64_t w[80];
        for (int i = 0; i < 16; i++) {
            uint64_t v = 0;
            for (int j = 0; j < 8; j++) v = (v << 8) | m[off + 8 * i + j];
            w[i] = v;
        }
        for (int i = 16; i < 80; i++)
            w[i] = w[i - 16] + (rotr(w[i - 15], 1) ^ rotr(w[i - 15], 8) ^ (w[i - 15] >> 7)) + w[i - 7] + (rotr(w[i - 2], 19) ^ rotr(w[i - 2], 61) ^ (w[i - 2] >> 6));
        uint64_t v[8];
        memcpy(v, h, sizeof v);
        for (int i = 0; i < 80; i++) {
            uint64_t t1 = v[7] + (rotr(v[4], 14) ^ rotr(v[4], 18) ^ rotr(v[4], 41)) + ((v[4] & v[5]) ^ (~v[4] & v[6])) + T.K[i] + w[i];
            uint64_t t2 = (rotr(v[0], 28) ^ rotr(v[0], 34) ^ rotr(v[0], 39)) + ((v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]));
            for (int k = 7; k > 0; k--) v[k] = v[k - 1];
            v[4] += t1;
            v[0] = t1 + t2;
        }
        for (int i = 0; i < 8; i++) h[i] += v[i];
    }
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(h[i] >> (56 - 8 * j));
}

// ---------------------------------------------------------------- STROBE-128 (the subset Merlin uses)
class strobe128 {
  public:
    explicit strobe128(const char *protocol_label) {
        memset(st_, 0, sizeof st_);
        uint8_t *s = bytes();
        static const uint8_t init[18] = {1, 168, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
        memcpy(s, init, 18);
        keccak_f1600(st_);
        pos_ = 0; pos_begin_ = 0; cur_flags_ = 0;
        meta_ad(protocol_label, strlen(protocol_label), false);
    }
    void meta_ad(const void *d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb((const uint8_t *)d, n); }
    void ad(const void *d, size_t n, bool more) { begin_op(FLAG_A, more); absorb((const uint8_t *)d, n); }
    void prf(void *d, size_t n, bool more) { begin_op(FLAG_I | FLAG_A | FLAG_C, more); squeeze((uint8_t *)d, n); }
    void key(const void *d, size_t n, bool more) { begin_op(FLAG_A | FLAG_C, more); overwrite((const uint8_t *)d, n); }
    // hand-off to / from the device-side continuation (rng_kernels.cuh): 200 B state, pos, pos_begin, cur_flags
    void export_state(uint8_t out[208]) const {
        memcpy(out, st_, 200);
        out[200] = pos_; out[201] = pos_begin_; out[202] = cur_flags_;
        memset(out + 203, 0, 5);
    }
    void import_state(const uint8_t in[208]) {
        memcpy(st_, in, 200);
        pos_ = in[200]; pos_begin_ = in[201]; cur_flags_ = in[202];
    }

  private:
    static const int R = 166;
    enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
    uint64_t st_[25];
    uint8_t pos_, pos_begin_, cur_flags_;
    uint8_t *bytes() { return (uint8_t *)st_; }
    void run_f() {
        uint8_t *s = bytes();
        s[pos_] ^= pos_begin_;
        s[pos_ + 1] ^= 0x04;
        s[R + 1] ^= 0x80;
        keccak_f1600(st_);
        pos_ = 0; pos_begin_ = 0;
    }
    void absorb(const uint8_t *d, size_t n) {
        uint8_t *s = bytes();
        for (size_t i = 0; i < n; i++) { s[pos_++] ^= d[i]; if (pos_ == R) run_f(); }
    }
    void overwrite(const uint8_t *d, size_t n) {
        uint8_t *s = bytes();
        for (size_t i = 0; i < n; i++) { s[pos_++] = d[i]; if (pos_ == R) run_f(); }
    }
    void squeeze(uint8_t *d, size_t n) {
        uint8_t *s = bytes();
        for (size_t i = 0; i < n; i++) { d[i] = s[pos_]; s[pos_++] = 0; if (pos_ == R) run_f(); }
    }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;
        uint8_t old_begin = pos_begin_;
        pos_begin_ = (uint8_t)(pos_ + 1);
        cur_flags_ = flags;
        uint8_t hdr[2] = {old_begin, flags};
        absorb(hdr, 2);
        if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
    }
};

inline void put_le32(uint8_t out[4], uint32_t x) { out[0] = (uint8_t)x; out[1] = (uint8_t)(x >> 8); out[2] = (uint8_t)(x >> 16); out[3] = (uint8_t)(x >> 24); }

// merlin::TranscriptRng
class merlin_rng {
  public:
    explicit merlin_rng(const strobe128 &s) : s_(s) {}
    void fill_bytes(uint8_t *dest, size_t n) {
        uint8_t len[4];
        put_le32(len, (uint32_t)n);
        s_.meta_ad(len, 4, false);
        s_.prf(dest, n, false);
    }
    sc random_scalar() {
        uint8_t b[64];
        fill_bytes(b, 64);
        return sc_from_wide(b);
    }
    void export_state(uint8_t out[208]) const { s_.export_state(out); }
    void import_state(const uint8_t in[208]) { s_.import_state(in); }
  private:
    strobe128 s_;
};
// merlin::TranscriptRngBuilder
class merlin_rng_builder {
  public:
    explicit merlin_rng_builder(const strobe128 &s) : s_(s) {}
    void rekey_with_witness_bytes(const char *label, const uint8_t *w, size_t n) {
        uint8_t len[4];
        put_le32(len, (uint32_t)n);
        s_.meta_ad(label, strlen(label), false);
        s_.meta_ad(len, 4, true);
        s_.key(w, n, false);
    }
    // external32: the 32 bytes the reference draws from thread_rng inside finalize (RNG contract)
    merlin_rng finalize(const uint8_t external32[32]) {
        s_.meta_ad("rng", 3, false);
        s_.key(external32, 32, false);
        return merlin_rng(s_);
    }
  private:
    strobe128 s_;
};

// merlin::Transcript + bulletproofs' TranscriptProtocol extension trait
class merlin_transcript {
  public:
    explicit merlin_transcript(const char *label) : s_("Merlin v1.0") { append_message("dom-sep", label, strlen(label)); }
    // labels and messages as (pointer, length) pairs: what the C ABI's bbp_transcript_* entry points receive
    merlin_transcript(const void *label, size_t label_len) : s_("Merlin v1.0") { append_message_l("dom-sep", 7, label, label_len); }
    void append_message_l(const void *label, size_t label_len, const void *msg, size_t n) {
        uint8_t len[4];
        put_le32(len, (uint32_t)n);
        s_.meta_ad(label, label_len, false);
        s_.meta_ad(len, 4, true);
        s_.ad(msg, n, false);
    }
    void challenge_bytes_l(const void *label, size_t label_len, uint8_t *dest, size_t n) {
        uint8_t len[4];
        put_le32(len, (uint32_t)n);
        s_.meta_ad(label, label_len, false);
        s_.meta_ad(len, 4, true);
        s_.prf(dest, n, false);
    }
    void append_message(const char *label, const void *msg, size_t n) {
        uint8_t len[4];
        put_le32(len, (uint32_t)n);
        s_.meta_ad(label, strlen(label), false);
        s_.meta_ad(len, 4, true);
        s_.ad(msg, n, false);
    }
    void append_u64(const char *label, uint64_t x) {
        uint8_t b[8];
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
        append_message(label, b, 8);
    }
    void challenge_bytes(const char *label, uint8_t *dest, size_t n) {
        uint8_t len[4];
        put_le32(len, (uint32_t)n);
        s_.meta_ad(label, strlen(label), false);
        s_.meta_ad(len, 4, true);
        s_.prf(dest, n, false);
    }
    merlin_rng_builder build_rng() const { return merlin_rng_builder(s_); }
    void export_state(uint8_t out[208]) const { s_.export_state(out); }   // hand-off to the device-side replay
    void import_state(const uint8_t in[208]) { s_.import_state(in); }

    void domain_sep(const char *name) { append_message("dom-sep", name, strlen(name)); }
    void r1cs_domain_sep() { domain_sep("r1cs v1"); }
    void r1cs_1phase_domain_sep() { domain_sep("r1cs-1phase"); }
    void innerproduct_domain_sep(uint64_t n) { domain_sep("ipp v1"); append_u64("n", n); }
    void rangeproof_domain_sep(uint64_t n, uint64_t m) { domain_sep("rangeproof v1"); append_u64("n", n); append_u64("m", m); }
    void append_scalar(const char *label, const sc &x) {
        uint8_t b[32];
        sc_tobytes(b, x);
        append_message(label, b, 32);
    }
    void append_point(const char *label, const uint8_t compressed[32]) { append_message(label, compressed, 32); }
    bool validate_and_append_point(const char *label, const uint8_t compressed[32]) {
        uint8_t acc = 0;
        for (int i = 0; i < 32; i++) acc |= compressed[i];
        if (!acc) return false;   // identity is rejected
        append_message(label, compressed, 32);
        return true;
    }
    sc challenge_scalar(const char *label) {
        uint8_t b[64];
        challenge_bytes(label, b, 64);
        return sc_from_wide(b);
    }
  private:
    strobe128 s_;
};

}  // namespace bbp
