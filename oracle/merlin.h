// ORACLE — test infrastructure only (see fe.h header).
// Merlin 1.3.0 transcripts (Cargo.lock:399-400): STROBE-128 over Keccak-f[1600], Transcript and
// TranscriptRng, plus the bulletproofs TranscriptProtocol helpers.
// Pinned by merlin's published test vector ("test protocol" / "some label" / "some data").
#pragma once
#include "hash.h"
#include "sc.h"
#include <string>

namespace orc {

struct strobe128 {
    static const int R = 166;
    static const uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32;
    uint64_t st[25];
    uint8_t pos, pos_begin, cur_flags;

    uint8_t *bytes() { return (uint8_t *)st; }

    explicit strobe128(const char *protocol_label) {
        memset(st, 0, sizeof(st));
        uint8_t *s = bytes();
        const uint8_t hdr[6] = {1, R + 2, 1, 0, 1, 96};
        memcpy(s, hdr, 6);
        memcpy(s + 6, "STROBEv1.0.2", 12);
        keccak_f1600(st);
        pos = 0; pos_begin = 0; cur_flags = 0;
        meta_ad((const uint8_t *)protocol_label, strlen(protocol_label), false);
    }
    void run_f() {
        uint8_t *s = bytes();
        s[pos] ^= pos_begin;
        s[pos + 1] ^= 0x04;
        s[R + 1] ^= 0x80;
        keccak_f1600(st);
        pos = 0; pos_begin = 0;
    }
    void absorb(const uint8_t *d, size_t n) {
        uint8_t *s = bytes();
        for (size_t i = 0; i < n; i++) { s[pos++] ^= d[i]; if (pos == R) run_f(); }
    }
    void overwrite(const uint8_t *d, size_t n) {
        uint8_t *s = bytes();
        for (size_t i = 0; i < n; i++) { s[pos++] = d[i]; if (pos == R) run_f(); }
    }
    void squeeze(uint8_t *d, size_t n) {
        uint8_t *s = bytes();
        for (size_t i = 0; i < n; i++) { d[i] = s[pos]; s[pos++] = 0; if (pos == R) run_f(); }
    }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;  // continuing the current operation (flags must match cur_flags)
        uint8_t old_begin = pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        uint8_t hdr[2] = {old_begin, flags};
        absorb(hdr, 2);
        bool force_f = (flags & (FLAG_C | FLAG_K)) != 0;
        if (force_f && pos != 0) run_f();
    }
    void meta_ad(const uint8_t *d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
    void ad(const uint8_t *d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
    void prf(uint8_t *d, size_t n, bool more) { begin_op(FLAG_I | FLAG_A | FLAG_C, more); squeeze(d, n); }
    void key(const uint8_t *d, size_t n, bool more) { begin_op(FLAG_A | FLAG_C, more); overwrite(d, n); }
};

static inline void le32(uint8_t out[4], uint32_t x) {
    out[0] = (uint8_t)x; out[1] = (uint8_t)(x >> 8); out[2] = (uint8_t)(x >> 16); out[3] = (uint8_t)(x >> 24);
}

struct transcript_rng {
    strobe128 s;
    explicit transcript_rng(const strobe128 &st) : s(st) {}
    void fill_bytes(uint8_t *dest, size_t n) {
        uint8_t len[4];
        le32(len, (uint32_t)n);
        s.meta_ad(len, 4, false);
        s.prf(dest, n, false);
    }
    sc random_scalar() {  // Scalar::random: 64 bytes, wide reduction
        uint8_t b[64];
        fill_bytes(b, 64);
        return sc_from_wide(b);
    }
};

struct transcript_rng_builder {
    strobe128 s;
    explicit transcript_rng_builder(const strobe128 &st) : s(st) {}
    void rekey_with_witness_bytes(const char *label, const uint8_t *w, size_t n) {
        uint8_t len[4];
        le32(len, (uint32_t)n);
        s.meta_ad((const uint8_t *)label, strlen(label), false);
        s.meta_ad(len, 4, true);
        s.key(w, n, false);
    }
    // `external32` stands in for the 32 bytes the reference draws from thread_rng (SURVEY.md §8b RNG contract)
    transcript_rng finalize(const uint8_t external32[32]) {
        s.meta_ad((const uint8_t *)"rng", 3, false);
        s.key(external32, 32, false);
        return transcript_rng(s);
    }
};

struct transcript {
    strobe128 s;
    explicit transcript(const char *label) : s("Merlin v1.0") {
        append_message("dom-sep", (const uint8_t *)label, strlen(label));
    }
    void append_message(const char *label, const uint8_t *msg, size_t n) {
        uint8_t len[4];
        le32(len, (uint32_t)n);
        s.meta_ad((const uint8_t *)label, strlen(label), false);
        s.meta_ad(len, 4, true);
        s.ad(msg, n, false);
    }
    void append_u64(const char *label, uint64_t x) {
        uint8_t b[8];
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
        append_message(label, b, 8);
    }
    void challenge_bytes(const char *label, uint8_t *dest, size_t n) {
        uint8_t len[4];
        le32(len, (uint32_t)n);
        s.meta_ad((const uint8_t *)label, strlen(label), false);
        s.meta_ad(len, 4, true);
        s.prf(dest, n, false);
    }
    transcript_rng_builder build_rng() const { return transcript_rng_builder(s); }

    // ---- bulletproofs TranscriptProtocol (transcript.rs of bulletproofs@4a05305, [UP]) ----
    void domain_sep(const char *name) { append_message("dom-sep", (const uint8_t *)name, strlen(name)); }
    void rangeproof_domain_sep(uint64_t n, uint64_t m) { domain_sep("rangeproof v1"); append_u64("n", n); append_u64("m", m); }
    void innerproduct_domain_sep(uint64_t n) { domain_sep("ipp v1"); append_u64("n", n); }
    void r1cs_domain_sep() { domain_sep("r1cs v1"); }
    void r1cs_1phase_domain_sep() { domain_sep("r1cs-1phase"); }
    void append_scalar(const char *label, const sc &x) {
        uint8_t b[32];
        sc_tobytes(b, x);
        append_message(label, b, 32);
    }
    void append_point(const char *label, const uint8_t compressed[32]) { append_message(label, compressed, 32); }
    // rejects the identity encoding (32 zero bytes)
    bool validate_and_append_point(const char *label, const uint8_t compressed[32]) {
        uint8_t acc = 0;
        for (int i = 0; i < 32; i++) acc |= compressed[i];
        if (acc == 0) return false;
        append_message(label, compressed, 32);
        return true;
    }
    sc challenge_scalar(const char *label) {
        uint8_t b[64];
        challenge_bytes(label, b, 64);
        return sc_from_wide(b);
    }
};

}  // namespace orc
