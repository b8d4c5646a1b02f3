// ORACLE — test infrastructure only (see fe.h header).
// Keccak-f[1600], SHAKE256, SHA3-512 (sha3 0.8.2 / keccak 0.1.0, Cargo.lock:648-649,366-367) and
// SHA-512 (sha2 0.8.0, Cargo.lock:636-637; used at src/blindbid/mod.rs:11,18).
// Checked against hashlib in tests/test_oracle_primitives.py.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

static inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

struct keccak_tables {
    uint64_t rc[24];
    int rot[25];   // rotation offset of lane x+5y
};
static inline keccak_tables keccak_make_tables() {
    keccak_tables T;
    // round constants from the degree-8 LFSR of FIPS 202 §3.2.5
    uint8_t lfsr = 1;
    for (int r = 0; r < 24; r++) {
        uint64_t c = 0;
        for (int j = 0; j < 7; j++) {
            if (lfsr & 1) c ^= 1ULL << ((1 << j) - 1);
            lfsr = (lfsr & 0x80) ? (uint8_t)((lfsr << 1) ^ 0x71) : (uint8_t)(lfsr << 1);
        }
        T.rc[r] = c;
    }
    // rho offsets: (x,y) walk of FIPS 202 §3.2.2
    for (int i = 0; i < 25; i++) T.rot[i] = 0;
    int x = 1, y = 0;
    for (int t = 0; t < 24; t++) {
        T.rot[x + 5 * y] = ((t + 1) * (t + 2) / 2) % 64;
        int nx = y, ny = (2 * x + 3 * y) % 5;
        x = nx; y = ny;
    }
    return T;
}
static inline const keccak_tables &keccak_tab() {
    static const keccak_tables T = keccak_make_tables();
    return T;
}

static inline void keccak_f1600(uint64_t A[25]) {
    const keccak_tables &T = keccak_tab();
    for (int round = 0; round < 24; round++) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) A[i] ^= D[i % 5];
        // rho + pi: B[y, 2x+3y] = rot(A[x,y])
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(A[x + 5 * y], T.rot[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) A[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        A[0] ^= T.rc[round];
    }
}

// sponge with byte-addressed little-endian state (host is little-endian)
struct sponge {
    uint64_t st[25];
    size_t rate, pos;
    uint8_t dsuffix;
    bool squeezing;
    sponge(size_t rate_bytes, uint8_t suffix) : rate(rate_bytes), pos(0), dsuffix(suffix), squeezing(false) {
        memset(st, 0, sizeof(st));
    }
    void absorb(const uint8_t *in, size_t n) {
        uint8_t *s = (uint8_t *)st;
        for (size_t i = 0; i < n; i++) {
            s[pos++] ^= in[i];
            if (pos == rate) { keccak_f1600(st); pos = 0; }
        }
    }
    void finish() {
        uint8_t *s = (uint8_t *)st;
        s[pos] ^= dsuffix;
        s[rate - 1] ^= 0x80;
        keccak_f1600(st);
        pos = 0;
        squeezing = true;
    }
    void squeeze(uint8_t *out, size_t n) {
        if (!squeezing) finish();
        uint8_t *s = (uint8_t *)st;
        for (size_t i = 0; i < n; i++) {
            if (pos == rate) { keccak_f1600(st); pos = 0; }
            out[i] = s[pos++];
        }
    }
};
struct shake256 : sponge { shake256() : sponge(136, 0x1f) {} };

static inline void sha3_512(uint8_t out[64], const uint8_t *in, size_t n) {
    sponge s(72, 0x06);
    s.absorb(in, n);
    s.squeeze(out, 64);
}

#include "sha512_k.inc"

static inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

static inline void sha512(uint8_t out[64], const uint8_t *in, size_t n) {
    uint64_t h[8];
    memcpy(h, SHA512_H0, sizeof(h));
    std::vector<uint8_t> msg(in, in + n);
    msg.push_back(0x80);
    while (msg.size() % 128 != 112) msg.push_back(0);
    for (int i = 0; i < 8; i++) msg.push_back(0);  // high 64 bits of the 128-bit length
    uint64_t bits = (uint64_t)n * 8;
    for (int i = 7; i >= 0; i--) msg.push_back((uint8_t)(bits >> (8 * i)));
    for (size_t off = 0; off < msg.size(); off += 128) {
        uint64_t w[80];
        for (int i = 0; i < 16; i++) {
            uint64_t v = 0;
            for (int j = 0; j < 8; j++) v = (v << 8) | msg[off + 8 * i + j];
            w[i] = v;
        }
        for (int i = 16; i < 80; i++) {
            uint64_t s0 = rotr64(w[i - 15], 1) ^ rotr64(w[i - 15], 8) ^ (w[i - 15] >> 7);
            uint64_t s1 = rotr64(w[i - 2], 19) ^ rotr64(w[i - 2], 61) ^ (w[i - 2] >> 6);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 80; i++) {
            uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
            uint64_t ch = (e & f) ^ (~e & g);
            uint64_t t1 = hh + S1 + ch + SHA512_K[i] + w[i];
            uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
            uint64_t maj = (a & b) ^ (a & c) ^ (b & c);
            uint64_t t2 = S0 + maj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(h[i] >> (56 - 8 * j));
}

}  // namespace orc
