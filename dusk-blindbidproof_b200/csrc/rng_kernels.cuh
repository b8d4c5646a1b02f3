// merlin::TranscriptRng on the device: the prover draws 2 n1 + 8 scalars per proof from a STROBE-128 PRF (one
// Keccak-f[1600] per draw, inherently sequential per proof), which is the host-side hot loop of Prover::prove
// (SURVEY.md §3.4 "scalar side", §7 "TranscriptRng is inherently serial"). For batches the chain runs here instead: one
// thread per proof continues the transcript RNG from the exported STROBE state, writes s_L and s_R straight into the
// witness arrays in HBM (no H2D of 2 n1 scalars per proof) and hands the state back for the later T-blinding draws.
// Bit-exact with keccak.h's strobe128 / merlin_rng (the host path used for small batches).
#pragma once
#include <cstdint>
#include "sc25519.cuh"
#include "sc_kernels.cuh"

namespace bbp {

#ifndef BBP_KECCAK_UNROLL
#define BBP_KECCAK_UNROLL 1
#endif
#define BBP_STROBE_R 166
#define BBP_STROBE_STATE_BYTES 208   // 200 B Keccak state, pos, pos_begin, cur_flags, 5 B padding

__constant__ uint64_t KECCAK_RC_DEV[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static constexpr int keccak_unroll_n = BBP_KECCAK_UNROLL;
__device__ inline void keccak_f1600_dev(uint64_t *s) {
    uint64_t a00 = s[0], a01 = s[1], a02 = s[2], a03 = s[3], a04 = s[4], a05 = s[5], a06 = s[6], a07 = s[7], a08 = s[8], a09 = s[9], a10 = s[10],
             a11 = s[11], a12 = s[12], a13 = s[13], a14 = s[14], a15 = s[15], a16 = s[16], a17 = s[17], a18 = s[18], a19 = s[19], a20 = s[20],
             a21 = s[21], a22 = s[22], a23 = s[23], a24 = s[24];
#pragma unroll (keccak_unroll_n)
    for (int r = 0; r < 24; r++) {
        uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                 c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
        uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
        a00 ^= d0; a05 ^= d0; a10 ^= d0; a15 ^= d0; a20 ^= d0;
        a01 ^= d1; a06 ^= d1; a11 ^= d1; a16 ^= d1; a21 ^= d1;
        a02 ^= d2; a07 ^= d2; a12 ^= d2; a17 ^= d2; a22 ^= d2;
        a03 ^= d3; a08 ^= d3; a13 ^= d3; a18 ^= d3; a23 ^= d3;
        a04 ^= d4; a09 ^= d4; a14 ^= d4; a19 ^= d4; a24 ^= d4;
        uint64_t b00 = a00, b10 = rotl64(a01, 1), b20 = rotl64(a02, 62), b05 = rotl64(a03, 28), b15 = rotl64(a04, 27), b16 = rotl64(a05, 36),
                 b01 = rotl64(a06, 44), b11 = rotl64(a07, 6), b21 = rotl64(a08, 55), b06 = rotl64(a09, 20), b07 = rotl64(a10, 3), b17 = rotl64(a11, 10),
                 b02 = rotl64(a12, 43), b12 = rotl64(a13, 25), b22 = rotl64(a14, 39), b23 = rotl64(a15, 41), b08 = rotl64(a16, 45), b18 = rotl64(a17, 15),
                 b03 = rotl64(a18, 21), b13 = rotl64(a19, 8), b14 = rotl64(a20, 18), b24 = rotl64(a21, 2), b09 = rotl64(a22, 61), b19 = rotl64(a23, 56),
                 b04 = rotl64(a24, 14);
        a00 = b00 ^ (~b01 & b02); a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
        a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
        a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
        a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
        a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
        a00 ^= KECCAK_RC_DEV[r];
    }
    s[0] = a00; s[1] = a01; s[2] = a02; s[3] = a03; s[4] = a04; s[5] = a05; s[6] = a06; s[7] = a07; s[8] = a08; s[9] = a09; s[10] = a10; s[11] = a11;
    s[12] = a12; s[13] = a13; s[14] = a14; s[15] = a15; s[16] = a16; s[17] = a17; s[18] = a18; s[19] = a19; s[20] = a20; s[21] = a21; s[22] = a22;
    s[23] = a23; s[24] = a24;
}

struct strobe_dev {
    uint64_t st[25];
    uint32_t pos, pos_begin;
    __device__ uint8_t *bytes() { return (uint8_t *)st; }
    // The permutation and the two composite operations are real calls: inlined, the 50-odd transcript operations of the
    // verifier replay grew k_verify_transcript to 131 k SASS instructions and a third of its issue slots went to
    // instruction-cache misses (one warp per SM has nothing to hide them behind).
    __device__ __noinline__ void run_f() {
        uint8_t *s = bytes();
        s[pos] ^= (uint8_t)pos_begin;
        s[pos + 1] ^= 0x04;
        s[BBP_STROBE_R + 1] ^= 0x80;
        keccak_f1600_dev(st);
        pos = 0; pos_begin = 0;
    }
    __device__ void absorb_byte(uint8_t b) {
        bytes()[pos++] ^= b;
        if (pos == BBP_STROBE_R) run_f();
    }
    // begin_op for a fresh (more = false) operation with the given flags
    __device__ void begin_op(uint8_t flags) {
        uint8_t old_begin = (uint8_t)pos_begin;
        pos_begin = pos + 1;
        absorb_byte(old_begin);
        absorb_byte(flags);
        if ((flags & (4 | 32)) && pos != 0) run_f();   // FLAG_C | FLAG_K
    }
    // ---- the Merlin subset of STROBE operations (bit-exact twins of keccak.h's strobe128 / merlin_transcript)
    __device__ void absorb(const uint8_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) absorb_byte(d[i]); }
    __device__ void meta_ad_label(const char *label, uint32_t n) { begin_op(16 | 2); for (uint32_t i = 0; i < n; i++) absorb_byte((uint8_t)label[i]); }
    __device__ void meta_ad_len(uint32_t len) {   // meta_ad(LE32(len), more = true): continues the running meta_ad
        absorb_byte((uint8_t)len); absorb_byte((uint8_t)(len >> 8)); absorb_byte((uint8_t)(len >> 16)); absorb_byte((uint8_t)(len >> 24));
    }
    __device__ __noinline__ void append_message(const char *label, uint32_t llen, const uint8_t *msg, uint32_t n) {
        meta_ad_label(label, llen);
        meta_ad_len(n);
        begin_op(2);                        // ad
        absorb(msg, n);
    }
    __device__ void append_u64(const char *label, uint32_t llen, uint64_t x) {
        uint8_t b[8];
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
        append_message(label, llen, b, 8);
    }
    __device__ void prf(uint8_t *out, uint32_t n) {
        begin_op(1 | 2 | 4);
        uint8_t *s = bytes();
        for (uint32_t i = 0; i < n; i++) {
            out[i] = s[pos];
            s[pos++] = 0;
            if (pos == BBP_STROBE_R) run_f();
        }
    }
    __device__ void key(const uint8_t *d, uint32_t n) {
        begin_op(2 | 4);                    // FLAG_A | FLAG_C
        uint8_t *s = bytes();
        for (uint32_t i = 0; i < n; i++) {
            s[pos++] = d[i];
            if (pos == BBP_STROBE_R) run_f();
        }
    }
    // TranscriptProtocol::challenge_scalar: 64 challenge bytes, wide-reduced
    __device__ __noinline__ sc challenge_scalar(const char *label, uint32_t llen) {
        meta_ad_label(label, llen);
        meta_ad_len(64);
        uint8_t b[64];
        prf(b, 64);
        return sc_from_wide(b);
    }
    // TranscriptRng::fill_bytes(64): meta_ad(LE32(64)) ; prf(64)
    __device__ void fill64(uint8_t *out) {
        begin_op(16 | 2);                   // FLAG_M | FLAG_A
        absorb_byte(64); absorb_byte(0); absorb_byte(0); absorb_byte(0);
        begin_op(1 | 2 | 4);                // FLAG_I | FLAG_A | FLAG_C
        uint8_t *s = bytes();
        for (int i = 0; i < 64; i++) {
            out[i] = s[pos];
            s[pos++] = 0;
            if (pos == BBP_STROBE_R) run_f();
        }
    }
};

// ---------------------------------------------------------------- warp-cooperative Keccak-f[1600] and STROBE-128
// One WARP advances one sponge: lane i < 25 holds state lane i = x + 5 y in two registers. A round is two gather steps by
// warp shuffles and a handful of logic instructions per thread:
//   theta   each lane fetches the five lanes of column x-1 and the five of column x+1 (10 64-bit shuffles), D = C[x-1] ^ rotl(C[x+1], 1)
//   rho     every lane rotates its own word by its own offset
//   pi+chi  lane (X, Y) fetches the rotated words that pi moves to (X, Y), (X+1, Y), (X+2, Y) (3 shuffles) and combines them
//   iota    lane 0
// against ~5000 dependent logic instructions of one thread walking all 25 lanes. The dependent path of a round is ~100
// clocks, and a batch of B sponges is B warps spread over all SMs instead of B / 32 (one thread per sponge left 16 warps on
// 148 SMs for 512 requests: a pure latency chain, 26 % of the prover's and 34 % of the batch verifier's GPU time).
__constant__ uint8_t KECCAK_RHO_LANE[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
struct keccak_warp {
    int m_base, p_base, s0, s1, s2;   // source lanes of the two gathers
    uint32_t rot;                     // rho offset of the lane's own word
    uint64_t iota_mask;               // all ones on lane 0
    __device__ __forceinline__ void init(uint32_t lane) {
        const uint32_t i = lane < 25 ? lane : lane - 25;   // the seven idle lanes shadow lanes 0..6 (their results are never read)
        const uint32_t x = i % 5, y = i / 5;
        m_base = (int)((x + 4) % 5); p_base = (int)((x + 1) % 5);
        const uint32_t x1 = (x + 1) % 5, x2 = (x + 2) % 5;
        s0 = (int)((x + 3 * y) % 5 + 5 * x); s1 = (int)((x1 + 3 * y) % 5 + 5 * x1); s2 = (int)((x2 + 3 * y) % 5 + 5 * x2);
        rot = KECCAK_RHO_LANE[i];
        iota_mask = lane == 0 ? ~0ull : 0ull;
    }
    __device__ __forceinline__ uint64_t rotl_own(uint64_t v) const {
        uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        if (rot & 32) { uint32_t t = lo; lo = hi; hi = t; }
        const uint32_t r = rot & 31;
        uint32_t nlo = __funnelshift_l(hi, lo, r), nhi = __funnelshift_l(lo, hi, r);
        return ((uint64_t)nhi << 32) | nlo;
    }
    __device__ __forceinline__ uint64_t permute(uint64_t a) const {
#pragma unroll 2
        for (int r = 0; r < 24; r++) {
            uint64_t cm = shfl64(a, m_base) ^ shfl64(a, m_base + 5) ^ shfl64(a, m_base + 10) ^ shfl64(a, m_base + 15) ^ shfl64(a, m_base + 20);
            uint64_t cp = shfl64(a, p_base) ^ shfl64(a, p_base + 5) ^ shfl64(a, p_base + 10) ^ shfl64(a, p_base + 15) ^ shfl64(a, p_base + 20);
            uint64_t c = rotl_own(a ^ cm ^ rotl64(cp, 1));
            uint64_t b0 = shfl64(c, s0), b1 = shfl64(c, s1), b2 = shfl64(c, s2);
            a = b0 ^ (~b1 & b2) ^ (KECCAK_RC_DEV[r] & iota_mask);
        }
        return a;
    }
};

// STROBE-128 with the sponge state of one warp in shared memory (200 bytes, 8-byte aligned, owned by the warp): absorb /
// overwrite / squeeze touch the bytes lane-parallel, the permutation runs on registers (keccak_warp). pos / pos_begin are
// warp-uniform. Bit-exact twin of strobe_dev and of keccak.h's strobe128.
struct strobe_warp {
    uint8_t *st;
    uint32_t pos, pos_begin, lane;
    keccak_warp K;
    __device__ __forceinline__ void attach(uint8_t *smem_state, uint32_t lane_) { st = smem_state; lane = lane_; K.init(lane_); }
    // 208-byte exported state (keccak.h export_state) from generic memory
    __device__ __forceinline__ void load(const uint8_t *src) {
        if (lane < 25) ((uint64_t *)st)[lane] = ((const uint64_t *)src)[lane];
        pos = src[200]; pos_begin = src[201];
        __syncwarp();
    }
    __device__ __forceinline__ void store(uint8_t *dst, uint8_t cur_flags) {
        __syncwarp();
        if (lane < 25) ((uint64_t *)dst)[lane] = ((const uint64_t *)st)[lane];
        if (lane == 0) { dst[200] = (uint8_t)pos; dst[201] = (uint8_t)pos_begin; dst[202] = cur_flags; }
    }
    __device__ __noinline__ void run_f() {
        if (lane == 0) { st[pos] ^= (uint8_t)pos_begin; st[pos + 1] ^= 0x04; st[BBP_STROBE_R + 1] ^= 0x80; }
        __syncwarp();
        uint64_t a = ((const uint64_t *)st)[lane < 25 ? lane : lane - 25];
        a = K.permute(a);
        if (lane < 25) ((uint64_t *)st)[lane] = a;
        __syncwarp();
        pos = 0; pos_begin = 0;
    }
    __device__ __forceinline__ void absorb_byte(uint8_t b) {
        if (lane == 0) st[pos] ^= b;
        __syncwarp();
        if (++pos == BBP_STROBE_R) run_f();
    }
    // up to 16 warp-uniform bytes held in (lo, hi), little-endian
    __device__ __forceinline__ void absorb_packed(uint64_t lo, uint64_t hi, uint32_t n) {
        if (pos + n < BBP_STROBE_R) {
            if (lane < n) st[pos + lane] ^= (uint8_t)((lane < 8 ? lo >> (8 * lane) : hi >> (8 * (lane - 8))) & 0xff);
            __syncwarp();
            pos += n;
            return;
        }
        for (uint32_t i = 0; i < n; i++) absorb_byte((uint8_t)((i < 8 ? lo >> (8 * i) : hi >> (8 * (i - 8))) & 0xff));
    }
    __device__ __forceinline__ void absorb(const uint8_t *d, uint32_t n) {
        while (n) {
            const uint32_t chunk = min(n, BBP_STROBE_R - pos);
            for (uint32_t i = lane; i < chunk; i += 32) st[pos + i] ^= d[i];
            __syncwarp();
            pos += chunk; d += chunk; n -= chunk;
            if (pos == BBP_STROBE_R) run_f();
        }
    }
    __device__ __forceinline__ void begin_op(uint8_t flags) {
        const uint32_t old_begin = pos_begin;
        pos_begin = pos + 1;
        absorb_packed((uint64_t)old_begin | ((uint64_t)flags << 8), 0, 2);
        if ((flags & (4 | 32)) && pos != 0) run_f();   // FLAG_C | FLAG_K
    }
    __device__ __forceinline__ void meta_ad_label(const char *label, uint32_t n) { begin_op(16 | 2); absorb((const uint8_t *)label, n); }
    __device__ __forceinline__ void meta_ad_len(uint32_t len) { absorb_packed(len, 0, 4); }
    __device__ __noinline__ void append_message(const char *label, uint32_t llen, const uint8_t *msg, uint32_t n) {
        meta_ad_label(label, llen);
        meta_ad_len(n);
        begin_op(2);
        absorb(msg, n);
    }
    __device__ __forceinline__ void append_u64(const char *label, uint32_t llen, uint64_t x) {
        meta_ad_label(label, llen);
        meta_ad_len(8);
        begin_op(2);
        absorb_packed(x, 0, 8);
    }
    // prf(64): after begin_op with FLAG_C the position is 0, so the 64 bytes are state lanes 0..7; every lane gets all 16 words
    __device__ __forceinline__ void prf64(uint32_t w[16]) {
        begin_op(1 | 2 | 4);
        const uint32_t *s32 = (const uint32_t *)st;
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = s32[i];
        __syncwarp();
        if (lane < 16) ((uint32_t *)st)[lane] = 0;
        __syncwarp();
        pos = 64;
    }
    // key(32 bytes): begin_op(A | C) leaves pos = 0, then the bytes overwrite state bytes 0..31
    __device__ __forceinline__ void key32(const uint8_t *d) {
        begin_op(2 | 4);
        if (lane < 32) st[lane] = d[lane];
        __syncwarp();
        pos = 32;
    }
    __device__ __noinline__ sc challenge_scalar(const char *label, uint32_t llen) {
        meta_ad_label(label, llen);
        meta_ad_len(64);
        uint32_t w[16];
        prf64(w);
        return sc_from_wide_words(w);
    }
    // TranscriptRng::fill_bytes(64)
    __device__ __forceinline__ void fill64(uint32_t w[16]) {
        begin_op(16 | 2);
        absorb_packed(64, 0, 4);
        prf64(w);
    }
    // the same, lane i < 16 receiving output word i only
    __device__ __forceinline__ uint32_t fill64_word() {
        begin_op(16 | 2);
        absorb_packed(64, 0, 4);
        begin_op(1 | 2 | 4);
        const uint32_t v = ((const uint32_t *)st)[lane & 15];
        __syncwarp();
        if (lane < 16) ((uint32_t *)st)[lane] = 0;
        __syncwarp();
        pos = 64;
        return v;
    }
};

// Steady state of consecutive 64-byte draws: every draw starts at pos = 64, pos_begin = 0 (the previous prf squeezed lanes
// 0..7 after a permutation), so the STROBE framing of one draw is a fixed pattern of byte XORs:
//   bytes 64..71 (lane 8):  old_begin = 0, flags M|A = 0x12, LE32(64), old_begin = 65, flags I|A|C = 0x07
//   run_f at pos = 72:      byte 72 ^= pos_begin (71), byte 73 ^= 0x04 (lane 9), byte 167 ^= 0x80 (lane 20, top byte)
// then Keccak-f, output = lanes 0..7, which are zeroed. The whole draw runs on registers.
#define BBP_STROBE_DRAW_LANE8 0x0741000000401200ULL
#define BBP_STROBE_DRAW_LANE9 0x0000000000000447ULL
#define BBP_STROBE_DRAW_LANE20 0x8000000000000000ULL

// states: [n_proofs][208] (in / out). Draw d of proof p is written RAW (64 B, 16 words) to raw[(p * n_draws + d) * 16];
// k_wide_reduce turns the raw draws into scalars in parallel, so that the sequential chain carries nothing but Keccak.
// The generic (byte-wise) path runs only until the steady state pos = 64, pos_begin = 0 is reached — normally one
// draw — after which the state lives in registers.
__global__ void __launch_bounds__(32) k_rng_draws(uint8_t *__restrict__ states, uint32_t n_proofs, uint32_t n_draws, uint32_t *__restrict__ raw) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_proofs) return;
    strobe_dev S;
    const uint64_t *in = (const uint64_t *)(states + (size_t)p * BBP_STROBE_STATE_BYTES);
    for (int i = 0; i < 25; i++) S.st[i] = in[i];
    const uint8_t *tail = states + (size_t)p * BBP_STROBE_STATE_BYTES + 200;
    S.pos = tail[0]; S.pos_begin = tail[1];
    uint32_t *out = raw + (size_t)p * n_draws * 16;
    uint32_t d = 0;
    for (; d < n_draws && !(S.pos == 64 && S.pos_begin == 0); d++) {
        uint8_t buf[64];
        S.fill64(buf);
        for (int i = 0; i < 16; i++)
            out[(size_t)d * 16 + i] = (uint32_t)buf[4 * i] | ((uint32_t)buf[4 * i + 1] << 8) | ((uint32_t)buf[4 * i + 2] << 16) | ((uint32_t)buf[4 * i + 3] << 24);
    }
    if (d < n_draws) {
        uint64_t r[25];
#pragma unroll
        for (int i = 0; i < 25; i++) r[i] = S.st[i];
#pragma unroll 1
        for (; d < n_draws; d++) {
            r[8] ^= BBP_STROBE_DRAW_LANE8;
            r[9] ^= BBP_STROBE_DRAW_LANE9;
            r[20] ^= BBP_STROBE_DRAW_LANE20;
            keccak_f1600_dev(r);
            uint4 *o = (uint4 *)(out + (size_t)d * 16);
            o[0] = make_uint4((uint32_t)r[0], (uint32_t)(r[0] >> 32), (uint32_t)r[1], (uint32_t)(r[1] >> 32));
            o[1] = make_uint4((uint32_t)r[2], (uint32_t)(r[2] >> 32), (uint32_t)r[3], (uint32_t)(r[3] >> 32));
            o[2] = make_uint4((uint32_t)r[4], (uint32_t)(r[4] >> 32), (uint32_t)r[5], (uint32_t)(r[5] >> 32));
            o[3] = make_uint4((uint32_t)r[6], (uint32_t)(r[6] >> 32), (uint32_t)r[7], (uint32_t)(r[7] >> 32));
#pragma unroll
            for (int i = 0; i < 8; i++) r[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < 25; i++) S.st[i] = r[i];
    }
    uint64_t *o = (uint64_t *)(states + (size_t)p * BBP_STROBE_STATE_BYTES);
    for (int i = 0; i < 25; i++) o[i] = S.st[i];
    uint8_t *ot = states + (size_t)p * BBP_STROBE_STATE_BYTES + 200;
    ot[0] = (uint8_t)S.pos; ot[1] = (uint8_t)S.pos_begin; ot[2] = 1 | 2 | 4;   // cur_flags after a prf
}

// The same chain with one WARP per proof (keccak_warp): the generic path runs on the shared-memory sponge until the steady
// state is reached, then the state lives in the lanes' registers; lanes 0..7 store the 64 output bytes of a draw as one
// coalesced 64-byte row. 4 proofs per block.
__global__ void __launch_bounds__(128) k_rng_draws_warp(uint8_t *__restrict__ states, uint32_t n_proofs, uint32_t n_draws, uint32_t *__restrict__ raw) {
    __shared__ uint64_t sm_state[4][26];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.x * 4 + warp;
    if (p >= n_proofs) return;   // whole warps leave together
    strobe_warp S;
    S.attach((uint8_t *)sm_state[warp], lane);
    S.load(states + (size_t)p * BBP_STROBE_STATE_BYTES);
    uint32_t *out = raw + (size_t)p * n_draws * 16;
    uint32_t d = 0;
    for (; d < n_draws && !(S.pos == 64 && S.pos_begin == 0); d++) {
        const uint32_t v = S.fill64_word();
        if (lane < 16) out[(size_t)d * 16 + lane] = v;
    }
    if (d < n_draws) {
        uint64_t a = ((const uint64_t *)S.st)[lane < 25 ? lane : lane - 25];
        const uint64_t frame = lane == 8 ? BBP_STROBE_DRAW_LANE8 : lane == 9 ? BBP_STROBE_DRAW_LANE9 : lane == 20 ? BBP_STROBE_DRAW_LANE20 : 0ull;
        const bool is_out = lane < 8;
#pragma unroll 1
        for (; d < n_draws; d++) {
            a = S.K.permute(a ^ frame);
            if (is_out) {
                ((uint2 *)(out + (size_t)d * 16))[lane] = make_uint2((uint32_t)a, (uint32_t)(a >> 32));
                a = 0;
            }
        }
        __syncwarp();
        if (lane < 25) ((uint64_t *)S.st)[lane] = a;
        __syncwarp();
    }
    S.store(states + (size_t)p * BBP_STROBE_STATE_BYTES, 1 | 2 | 4);   // cur_flags after a prf
}

// Scalar::from_bytes_mod_order_wide over the raw draws: draw d of proof p -> out[(d / per_vec) * vec_stride + p * per_vec + d % per_vec]
// (per_vec = n1, vec_stride = n_proofs * n1: the first n1 draws fill s_L, the next n1 fill s_R)
__global__ void __launch_bounds__(128) k_wide_reduce(const uint32_t *__restrict__ raw, uint32_t n_proofs, uint32_t n_draws, uint32_t per_vec, size_t vec_stride,
                                                     sc *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_proofs * n_draws) return;
    uint32_t p = (uint32_t)(i / n_draws), d = (uint32_t)(i % n_draws);
    uint32_t w[16];
    const uint4 *q = (const uint4 *)(raw + i * 16);
#pragma unroll
    for (int k = 0; k < 4; k++) { uint4 v = q[k]; w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
    sc lo = sc_reduce_words(w);
    sc r2 = sc_r2();
    sc hi = sc_montmul(w + 8, r2.v);
    out[(size_t)(d / per_vec) * vec_stride + (size_t)p * per_vec + d % per_vec] = sc_add(lo, hi);
}

// ---------------------------------------------------------------- SHAKE256 draw stream on the device (aggregated range proofs)
// The range-proof prover draws 2 m (1 + n) + 2 m scalars per proof from SHAKE256 streams (rangeproof.cuh RNG contract):
// one thread per (proof, party) squeezes that party's stream SHAKE256(seed || v_blinding || LE64(value) || LE32(party))
// (76 absorbed bytes; one Keccak-f per 136 bytes, state in registers), writing raw 64-byte draws; k_rp_draw_scatter reduces them mod l in parallel and routes them to
// s_L / s_R or to the small per-party list the host sums (a_blinding, s_blinding, t_1 / t_2 blindings).
// raw: [n_proofs][m][per_party][16 words], per_party = 4 + 2 nbits.
// wkeys: [n_proofs][m] x 40 B = the party's v_blinding (32 B canonical) and value (LE64)
__global__ void __launch_bounds__(64) k_shake_draws(const uint8_t *__restrict__ seeds, const uint8_t *__restrict__ wkeys, uint32_t n_proofs, uint32_t m,
                                                    uint32_t per_party, uint32_t *__restrict__ raw) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_proofs * m) return;
    const uint32_t p = idx / m, j = idx % m;
    uint64_t r[25];
#pragma unroll
    for (int i = 0; i < 25; i++) r[i] = 0;
    const uint64_t *sd = (const uint64_t *)(seeds + 32 * (size_t)p);
    const uint64_t *wk = (const uint64_t *)(wkeys + 40 * (size_t)idx);
    r[0] = sd[0]; r[1] = sd[1]; r[2] = sd[2]; r[3] = sd[3];
    r[4] = wk[0]; r[5] = wk[1]; r[6] = wk[2]; r[7] = wk[3];   // v_blinding
    r[8] = wk[4];                                               // LE64(value)
    r[9] = (uint64_t)j | (0x1fULL << 32);  // LE32(party), then SHAKE domain separation + first pad bit after the 76 absorbed bytes
    r[16] ^= 0x8000000000000000ULL;        // last pad bit at byte 135 (rate 136)
    keccak_f1600_dev(r);
    uint64_t *out = (uint64_t *)(raw + (size_t)idx * per_party * 16);
    const size_t total_words = (size_t)per_party * 8;   // 64-bit words to produce
    size_t w = 0;
    while (w < total_words) {
#pragma unroll
        for (int i = 0; i < 17; i++)
            if (w + i < total_words) out[w + i] = r[i];
        w += 17;
        if (w < total_words) keccak_f1600_dev(r);
    }
}
// draw e of party j of proof p: 0 = a_blinding, 1 = s_blinding, then s_L[0..n), s_R[0..n), then t_1 / t_2 blindings.
// small: [n_proofs][4][m] = a, s, t1, t2 blindings.
__global__ void __launch_bounds__(128) k_rp_draw_scatter(const uint32_t *__restrict__ raw, uint32_t n_proofs, uint32_t m, uint32_t nbits, sc *__restrict__ sL,
                                                         sc *__restrict__ sR, sc *__restrict__ small) {
    const uint32_t per_party = 4 + 2 * nbits;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_proofs * m * per_party) return;
    uint32_t e = (uint32_t)(i % per_party), j = (uint32_t)((i / per_party) % m), p = (uint32_t)(i / ((size_t)per_party * m));
    uint32_t w[16];
    const uint4 *q = (const uint4 *)(raw + i * 16);
#pragma unroll
    for (int k = 0; k < 4; k++) { uint4 v = q[k]; w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
    sc lo = sc_reduce_words(w);
    sc r2 = sc_r2();
    sc hi = sc_montmul(w + 8, r2.v);
    sc val = sc_add(lo, hi);
    const size_t nm = (size_t)m * nbits;
    if (e == 0) small[((size_t)p * 4 + 0) * m + j] = val;
    else if (e == 1) small[((size_t)p * 4 + 1) * m + j] = val;
    else if (e < 2 + nbits) sL[(size_t)p * nm + (size_t)j * nbits + (e - 2)] = val;
    else if (e < 2 + 2 * nbits) sR[(size_t)p * nm + (size_t)j * nbits + (e - 2 - nbits)] = val;
    else small[((size_t)p * 4 + 2 + (e - 2 - 2 * nbits)) * m + j] = val;
}

// ---------------------------------------------------------------- the verifier's transcript replay on the device
// One thread per request replays Verifier::verify's Fiat-Shamir transcript from the request's blob
//   [V_0..V_{m-1} | A_I1 A_O1 S1 A_I2 A_O2 S2 | T_1 T_3 T_4 T_5 T_6 | L_0 R_0 .. | t_x t_x_blinding e_blinding | a b]   (32 B each)
// starting from the exported state after Transcript::new(label) + "r1cs v1" (identical for every request), derives
// y, z, u, x, w, the u_j, the verifier's random scalar r (TranscriptRng keyed with the request's rng seed), inverts y and
// the u_j with one exponentiation, and writes the request's challenge block and the transcript-dependent dynamic scalars.
// The identity / canonical checks of validate_and_append_point and from_bytes are done on the host while parsing.
#define BBP_LBL(s) s, (uint32_t)(sizeof(s) - 1)
struct transcript_init { uint8_t state[BBP_STROBE_STATE_BYTES]; };

__global__ void __launch_bounds__(32) k_verify_transcript(transcript_init init, const uint8_t *__restrict__ blobs, uint32_t blob_stride, const uint8_t *__restrict__ seeds,
                                                          uint32_t n_req, uint32_t m, uint32_t lg, uint64_t n_ipp, sc *__restrict__ chal, sc *__restrict__ dyn,
                                                          uint32_t dyn_stride) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_req) return;
    strobe_dev S;
    {
        const uint64_t *in = (const uint64_t *)init.state;
        for (int i = 0; i < 25; i++) S.st[i] = in[i];
        S.pos = init.state[200]; S.pos_begin = init.state[201];
    }
    const uint8_t *blob = blobs + (size_t)p * blob_stride;
    const uint8_t *pts = blob + 32 * (size_t)m;                   // A_I1 ...
    const uint8_t *lr = pts + 32 * 11;
    const uint8_t *scal = lr + 64 * (size_t)lg;                   // t_x, t_x_blinding, e_blinding, a, b
    sc *c = chal + (size_t)p * CH_N;
    for (uint32_t i = 0; i < m; i++) S.append_message(BBP_LBL("V"), blob + 32 * (size_t)i, 32);
    S.append_u64(BBP_LBL("m"), m);
    S.append_message(BBP_LBL("A_I1"), pts, 32);
    S.append_message(BBP_LBL("A_O1"), pts + 32, 32);
    S.append_message(BBP_LBL("S1"), pts + 64, 32);
    S.append_message(BBP_LBL("dom-sep"), (const uint8_t *)"r1cs-1phase", 11);
    S.append_message(BBP_LBL("A_I2"), pts + 96, 32);
    S.append_message(BBP_LBL("A_O2"), pts + 128, 32);
    S.append_message(BBP_LBL("S2"), pts + 160, 32);
    sc y = S.challenge_scalar(BBP_LBL("y"));
    sc z = S.challenge_scalar(BBP_LBL("z"));
    S.append_message(BBP_LBL("T_1"), pts + 192, 32);
    S.append_message(BBP_LBL("T_3"), pts + 224, 32);
    S.append_message(BBP_LBL("T_4"), pts + 256, 32);
    S.append_message(BBP_LBL("T_5"), pts + 288, 32);
    S.append_message(BBP_LBL("T_6"), pts + 320, 32);
    sc u = S.challenge_scalar(BBP_LBL("u"));
    sc x = S.challenge_scalar(BBP_LBL("x"));
    S.append_message(BBP_LBL("t_x"), scal, 32);
    S.append_message(BBP_LBL("t_x_blinding"), scal + 32, 32);
    S.append_message(BBP_LBL("e_blinding"), scal + 64, 32);
    sc w = S.challenge_scalar(BBP_LBL("w"));
    S.append_message(BBP_LBL("dom-sep"), (const uint8_t *)"ipp v1", 6);
    S.append_u64(BBP_LBL("n"), n_ipp);
    // prefix products for the batch inversion of (u_0 .. u_{lg-1}, y), in the Montgomery domain
    sc acc = sc_to_mont(sc_one());
    for (uint32_t j = 0; j < lg; j++) {
        S.append_message(BBP_LBL("L"), lr + 64 * (size_t)j, 32);
        S.append_message(BBP_LBL("R"), lr + 64 * (size_t)j + 32, 32);
        sc uj = S.challenge_scalar(BBP_LBL("u"));
        c[CH_UJ0 + j] = uj;
        c[CH_UJ0 + lg + j] = acc;          // prefix (Montgomery form), replaced by the inverse below
        acc = sc_montmul(acc.v, sc_to_mont(uj).v);
    }
    sc pre_y = acc;
    acc = sc_montmul(acc.v, sc_to_mont(y).v);
    // verifier randomness: transcript.build_rng().finalize(rng_seed) -> one scalar
    S.meta_ad_label(BBP_LBL("rng"));
    S.key(seeds + 32 * (size_t)p, 32);
    uint8_t rb[64];
    S.fill64(rb);
    sc r = sc_from_wide(rb);
    // acc = (prod u_j * y) R ; invert once: x^(l-2) in the Montgomery domain
    sc inv = sc_to_mont(sc_one());
    for (int i = 252; i >= 0; i--) {
        inv = sc_montmul(inv.v, inv.v);
        uint32_t e = sc_l_limb(i >> 5);
        if (i < 32) e = sc_l_limb(0) - 2;
        if ((e >> (i & 31)) & 1) inv = sc_montmul(inv.v, acc.v);
    }
    sc yinv = sc_from_mont(sc_montmul(inv.v, pre_y.v));
    inv = sc_montmul(inv.v, sc_to_mont(y).v);
    sc *d = dyn + (size_t)p * dyn_stride + m;
    for (uint32_t j = lg; j-- > 0;) {
        sc uj = c[CH_UJ0 + j];
        sc ujinv = sc_from_mont(sc_montmul(inv.v, c[CH_UJ0 + lg + j].v));
        inv = sc_montmul(inv.v, sc_to_mont(uj).v);
        c[CH_UJ0 + lg + j] = ujinv;
        d[11 + 2 * j] = sc_mul(uj, uj);            // weight of L_j
        d[11 + 2 * j + 1] = sc_mul(ujinv, ujinv);  // weight of R_j
    }
    sc xx = sc_mul(x, x), xxx = sc_mul(xx, x), rxx = sc_mul(r, xx);
    d[0] = x; d[1] = xx; d[2] = xxx; d[3] = sc_mul(u, x); d[4] = sc_mul(u, xx); d[5] = sc_mul(u, xxx);
    d[6] = sc_mul(r, x); d[7] = sc_mul(rxx, x); d[8] = sc_mul(rxx, xx); d[9] = sc_mul(rxx, xxx); d[10] = sc_mul(sc_mul(rxx, xx), xx);
    sc t;
    c[CH_Y] = y; c[CH_YINV] = yinv; c[CH_Z] = z; c[CH_X] = x; c[CH_U] = u; c[CH_W] = w; c[CH_R] = r;
    const uint32_t *sw = (const uint32_t *)scal;
    for (int k = 0; k < 8; k++) t.v[k] = sw[k];
    c[CH_TX] = t;
    for (int k = 0; k < 8; k++) t.v[k] = sw[8 + k];
    c[CH_TXBL] = t;
    for (int k = 0; k < 8; k++) t.v[k] = sw[16 + k];
    c[CH_EBL] = t;
    for (int k = 0; k < 8; k++) t.v[k] = sw[24 + k];
    c[CH_A] = t;
    for (int k = 0; k < 8; k++) t.v[k] = sw[32 + k];
    c[CH_B] = t;
    c[CH_RHO] = sc_one();
}

// The same replay with one WARP per request (strobe_warp / keccak_warp): the ~50 permutations and ~60 transcript operations
// of a request run lane-parallel; the scalar arithmetic that follows (challenge reductions, one batched inversion, the
// dynamic-point weights) is computed redundantly by every lane (uniform control flow) and stored by lane 0. 4 requests per block.
// x^-1 in the Montgomery domain (x^(l-2), square and multiply over the constant exponent)
__device__ __noinline__ sc sc_invert_mont(sc accM) {
    sc inv = sc_to_mont(sc_one());
#pragma unroll 1
    for (int i = 252; i >= 0; i--) {
        inv = mm(inv, inv);
        uint32_t e = sc_l_limb(i >> 5);
        if (i < 32) e = sc_l_limb(0) - 2;
        if ((e >> (i & 31)) & 1) inv = mm(inv, accM);
    }
    return inv;
}

// The same for PUBLIC values (the verifier's challenges): binary extended Euclid (HAC 14.61) on 256-bit integers — about
// 350 halvings and 180 subtractions of 8 limbs (~19 k instructions) instead of 265 Montgomery products (~61 k). Variable
// time, which the verifier may be; every lane of the warp computes the same value, so the data-dependent loops do not
// diverge. Input and output in the Montgomery domain like sc_invert_mont; 0 -> 0 as x^(l-2) gives.
__device__ __forceinline__ void u256_shr1(uint32_t *a) {
#pragma unroll
    for (int i = 0; i < 7; i++) a[i] = __funnelshift_r(a[i], a[i + 1], 1);
    a[7] >>= 1;
}
// x / 2 mod l for x < l: (x + (x odd ? l : 0)) >> 1
__device__ __forceinline__ void sc_half(uint32_t *x) {
    const uint32_t m = 0u - (x[0] & 1u);
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+&r"(x[0]), "+&r"(x[1]), "+&r"(x[2]), "+&r"(x[3]), "+&r"(x[4]), "+&r"(x[5]), "+&r"(x[6]), "+&r"(x[7])
        : "r"(sc_l_limb(0) & m), "r"(sc_l_limb(1) & m), "r"(sc_l_limb(2) & m), "r"(sc_l_limb(3) & m), "r"(sc_l_limb(4) & m), "r"(sc_l_limb(5) & m),
          "r"(sc_l_limb(6) & m), "r"(sc_l_limb(7) & m));
    u256_shr1(x);
}
// t = a - b over 256 bits; returns the borrow (1 when a < b)
__device__ __forceinline__ uint32_t u256_sub(uint32_t *t, const uint32_t *a, const uint32_t *b) {
    uint32_t bw;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return bw & 1u;
}
__device__ __forceinline__ bool u256_is_one(const uint32_t *a) {
    return (a[0] ^ 1u | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0;
}
__device__ __noinline__ sc sc_invert_mont_vartime(sc accM) {
    sc a = sc_from_mont(accM);   // < l
    if (sc_iszero(a)) return a;
    uint32_t u[8], v[8], t[8];
    sc x1 = sc_one(), x2 = sc_zero();
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = a.v[i]; v[i] = sc_l_limb(i); }
#pragma unroll 1
    while (!u256_is_one(u) && !u256_is_one(v)) {
#pragma unroll 1
        while (!(u[0] & 1u)) { u256_shr1(u); sc_half(x1.v); }
#pragma unroll 1
        while (!(v[0] & 1u)) { u256_shr1(v); sc_half(x2.v); }
        if (!u256_sub(t, u, v)) {      // u >= v
#pragma unroll
            for (int i = 0; i < 8; i++) u[i] = t[i];
            x1 = sc_sub(x1, x2);
        } else {
            u256_sub(v, v, u);
            x2 = sc_sub(x2, x1);
        }
    }
    return sc_to_mont(u256_is_one(u) ? x1 : x2);
}

// phase 0: the whole replay in one launch. phase 1: up to the challenges y, z (plus y^-1) — everything k_powers needs — then the
// sponge is parked in `states` (208 B per request); phase 2: the rest, resumed from there. With the split the power tables
// are built on a second stream while the (latency-bound) remainder of the replay runs.
__global__ void __launch_bounds__(128) k_verify_transcript_warp(transcript_init init, const uint8_t *__restrict__ blobs, uint32_t blob_stride,
                                                                const uint8_t *__restrict__ seeds, uint32_t n_req, uint32_t m, uint32_t lg, uint64_t n_ipp,
                                                                sc *__restrict__ chal, sc *__restrict__ dyn, uint32_t dyn_stride, uint32_t phase,
                                                                uint8_t *__restrict__ states) {
    __shared__ uint64_t sm_state[4][26];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.x * 4 + warp;
    if (p >= n_req) return;   // whole warps leave together
    strobe_warp S;
    S.attach((uint8_t *)sm_state[warp], lane);
    const bool lead = lane == 0;
    const uint8_t *blob = blobs + (size_t)p * blob_stride;
    const uint8_t *pts = blob + 32 * (size_t)m;                   // A_I1 ...
    const uint8_t *lr = pts + 32 * 11;
    const uint8_t *scal = lr + 64 * (size_t)lg;                   // t_x, t_x_blinding, e_blinding, a, b
    sc *c = chal + (size_t)p * CH_N;
    sc y, z;
    if (phase != 2) {
        S.load(init.state);
        for (uint32_t i = 0; i < m; i++) S.append_message(BBP_LBL("V"), blob + 32 * (size_t)i, 32);
        S.append_u64(BBP_LBL("m"), m);
        S.append_message(BBP_LBL("A_I1"), pts, 32);
        S.append_message(BBP_LBL("A_O1"), pts + 32, 32);
        S.append_message(BBP_LBL("S1"), pts + 64, 32);
        S.append_message(BBP_LBL("dom-sep"), (const uint8_t *)"r1cs-1phase", 11);
        S.append_message(BBP_LBL("A_I2"), pts + 96, 32);
        S.append_message(BBP_LBL("A_O2"), pts + 128, 32);
        S.append_message(BBP_LBL("S2"), pts + 160, 32);
        y = S.challenge_scalar(BBP_LBL("y"));
        z = S.challenge_scalar(BBP_LBL("z"));
        if (phase == 1) {
            sc yinv = sc_from_mont(sc_invert_mont_vartime(sc_to_mont(y)));
            if (lead) { c[CH_Y] = y; c[CH_Z] = z; c[CH_YINV] = yinv; }
            S.store(states + (size_t)p * BBP_STROBE_STATE_BYTES, 0);
            return;
        }
    } else {
        S.load(states + (size_t)p * BBP_STROBE_STATE_BYTES);
        y = c[CH_Y]; z = c[CH_Z];
    }
    S.append_message(BBP_LBL("T_1"), pts + 192, 32);
    S.append_message(BBP_LBL("T_3"), pts + 224, 32);
    S.append_message(BBP_LBL("T_4"), pts + 256, 32);
    S.append_message(BBP_LBL("T_5"), pts + 288, 32);
    S.append_message(BBP_LBL("T_6"), pts + 320, 32);
    sc u = S.challenge_scalar(BBP_LBL("u"));
    sc x = S.challenge_scalar(BBP_LBL("x"));
    S.append_message(BBP_LBL("t_x"), scal, 32);
    S.append_message(BBP_LBL("t_x_blinding"), scal + 32, 32);
    S.append_message(BBP_LBL("e_blinding"), scal + 64, 32);
    sc w = S.challenge_scalar(BBP_LBL("w"));
    S.append_message(BBP_LBL("dom-sep"), (const uint8_t *)"ipp v1", 6);
    S.append_u64(BBP_LBL("n"), n_ipp);
    // prefix products for the batch inversion of (u_0 .. u_{lg-1}, y), in the Montgomery domain
    sc acc = sc_to_mont(sc_one());
    for (uint32_t j = 0; j < lg; j++) {
        S.append_message(BBP_LBL("L"), lr + 64 * (size_t)j, 32);
        S.append_message(BBP_LBL("R"), lr + 64 * (size_t)j + 32, 32);
        sc uj = S.challenge_scalar(BBP_LBL("u"));
        if (lead) {
            c[CH_UJ0 + j] = uj;
            c[CH_UJ0 + lg + j] = acc;      // prefix (Montgomery form), replaced by the inverse below
        }
        acc = mm(acc, sc_to_mont(uj));
    }
    __syncwarp();                          // lane 0's stores to c[] are read back by every lane below
    sc pre_y = acc;
    if (phase == 0) acc = mm(acc, sc_to_mont(y));   // phase 2: y^-1 was computed by phase 1, only the u_j are inverted here
    // verifier randomness: transcript.build_rng().finalize(rng_seed) -> one scalar
    S.meta_ad_label(BBP_LBL("rng"));
    S.key32(seeds + 32 * (size_t)p);
    uint32_t rw[16];
    S.fill64(rw);
    sc r = sc_from_wide_words(rw);
    // acc = (prod u_j [* y]) R ; invert once, in the Montgomery domain
    sc inv = sc_invert_mont_vartime(acc);   // public challenges: binary Euclid instead of x^(l-2)
    sc yinv;
    if (phase == 0) {
        yinv = sc_from_mont(mm(inv, pre_y));
        inv = mm(inv, sc_to_mont(y));
    } else {
        yinv = c[CH_YINV];
    }
    sc *d = dyn + (size_t)p * dyn_stride + m;
#pragma unroll 1
    for (uint32_t j = lg; j-- > 0;) {
        sc uj = c[CH_UJ0 + j];
        sc ujinv = sc_from_mont(mm(inv, c[CH_UJ0 + lg + j]));
        inv = mm(inv, sc_to_mont(uj));
        sc w_l = sc_mul(uj, uj), w_r = sc_mul(ujinv, ujinv);
        __syncwarp();                      // every lane has read c[CH_UJ0 + lg + j] before lane 0 overwrites it
        if (lead) {
            c[CH_UJ0 + lg + j] = ujinv;
            d[11 + 2 * j] = w_l;           // weight of L_j
            d[11 + 2 * j + 1] = w_r;       // weight of R_j
        }
    }
    if (!lead) return;
    sc xx = sc_mul(x, x), xxx = sc_mul(xx, x), rxx = sc_mul(r, xx);
    d[0] = x; d[1] = xx; d[2] = xxx; d[3] = sc_mul(u, x); d[4] = sc_mul(u, xx); d[5] = sc_mul(u, xxx);
    d[6] = sc_mul(r, x); d[7] = sc_mul(rxx, x); d[8] = sc_mul(rxx, xx); d[9] = sc_mul(rxx, xxx); d[10] = sc_mul(sc_mul(rxx, xx), xx);
    sc t;
    c[CH_Y] = y; c[CH_YINV] = yinv; c[CH_Z] = z; c[CH_X] = x; c[CH_U] = u; c[CH_W] = w; c[CH_R] = r;
    const uint32_t *sw = (const uint32_t *)scal;
    for (int k = 0; k < 8; k++) t.v[k] = sw[k];
    c[CH_TX] = t;
    for (int k = 0; k < 8; k++) t.v[k] = sw[8 + k];
    c[CH_TXBL] = t;
    for (int k = 0; k < 8; k++) t.v[k] = sw[16 + k];
    c[CH_EBL] = t;
    for (int k = 0; k < 8; k++) t.v[k] = sw[24 + k];
    c[CH_A] = t;
    for (int k = 0; k < 8; k++) t.v[k] = sw[32 + k];
    c[CH_B] = t;
    c[CH_RHO] = sc_one();
}


// ---------------------------------------------------------------- aggregated range proofs: verifier replay, one warp per proof
// RangeProof::verify_multiple's Fiat-Shamir replay and O(m + lg n) scalar work (rangeproof.cuh: rp_verify_group restates it
// on the host for small batches): challenges y, z, x, w, u_j, the verifier's two random scalars c and rho, one batched
// inversion of (u_j, y, y - 1, z - 1), delta(y, z), the B / B_blinding coefficients and the dynamic-point scalars. Everything
// is written UNWEIGHTED (chal0 / dyn0); k_rp_apply_weights applies the per-pass weight. blob = commitments (m x 32) | proof.
__global__ void __launch_bounds__(128) k_rp_verify_transcript_warp(transcript_init init, const uint8_t *__restrict__ blobs, uint32_t blob_stride,
                                                                   const uint8_t *__restrict__ seeds, uint32_t n_req, uint32_t m, uint32_t nbits, uint32_t lg,
                                                                   sc *__restrict__ chal0, sc *__restrict__ dyn0, uint32_t ds) {
    __shared__ uint64_t sm_state[4][26];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.x * 4 + warp;
    if (p >= n_req) return;   // whole warps leave together
    strobe_warp S;
    S.attach((uint8_t *)sm_state[warp], lane);
    S.load(init.state);
    const bool lead = lane == 0;
    const uint8_t *blob = blobs + (size_t)p * blob_stride;
    const uint8_t *pf = blob + 32 * (size_t)m;            // A S T_1 T_2 | t_x t_x_blinding e_blinding | L_0 R_0 .. | a b
    const uint8_t *LR = pf + 224;
    const uint64_t nm = (uint64_t)nbits * m;
    sc *c = chal0 + (size_t)p * CH_N;
    sc *d = dyn0 + (size_t)p * ds;
    for (uint32_t j = 0; j < m; j++) S.append_message(BBP_LBL("V"), blob + 32 * (size_t)j, 32);
    S.append_message(BBP_LBL("A"), pf, 32);
    S.append_message(BBP_LBL("S"), pf + 32, 32);
    const sc y = S.challenge_scalar(BBP_LBL("y"));
    const sc z = S.challenge_scalar(BBP_LBL("z"));
    S.append_message(BBP_LBL("T_1"), pf + 64, 32);
    S.append_message(BBP_LBL("T_2"), pf + 96, 32);
    const sc x = S.challenge_scalar(BBP_LBL("x"));
    S.append_message(BBP_LBL("t_x"), pf + 128, 32);
    S.append_message(BBP_LBL("t_x_blinding"), pf + 160, 32);
    S.append_message(BBP_LBL("e_blinding"), pf + 192, 32);
    const sc w = S.challenge_scalar(BBP_LBL("w"));
    S.append_message(BBP_LBL("dom-sep"), (const uint8_t *)"ipp v1", 6);
    S.append_u64(BBP_LBL("n"), nm);
    // prefix products for the batched inversion, in the Montgomery domain: u_0 .. u_{lg-1}, y, y - 1, z - 1
    sc acc = sc_to_mont(sc_one());
    for (uint32_t j = 0; j < lg; j++) {
        S.append_message(BBP_LBL("L"), LR + 64 * (size_t)j, 32);
        S.append_message(BBP_LBL("R"), LR + 64 * (size_t)j + 32, 32);
        sc uj = S.challenge_scalar(BBP_LBL("u"));
        if (lead) {
            c[CH_UJ0 + j] = uj;
            c[CH_UJ0 + lg + j] = acc;      // prefix (Montgomery form), replaced by the inverse below
        }
        acc = mm(acc, sc_to_mont(uj));
    }
    __syncwarp();
    // verifier randomness once every proof byte has been absorbed: c merges the two equations, rho weighs the request
    S.meta_ad_label(BBP_LBL("rng"));
    S.key32(seeds + 32 * (size_t)p);
    uint32_t rw[16];
    S.fill64(rw);
    const sc cc = sc_from_wide_words(rw);
    S.fill64(rw);
    const sc rho = sc_from_wide_words(rw);
    const sc one = sc_one();
    sc ym1 = sc_sub(y, one), zm1 = sc_sub(z, one);
    const bool y_is_1 = sc_iszero(ym1), z_is_1 = sc_iszero(zm1);
    if (y_is_1) ym1 = one;
    if (z_is_1) zm1 = one;
    const sc yM = sc_to_mont(y), ym1M = sc_to_mont(ym1), zm1M = sc_to_mont(zm1);
    const sc pre_y = acc;
    acc = mm(acc, yM);
    const sc pre_ym1 = acc;
    acc = mm(acc, ym1M);
    const sc pre_zm1 = acc;
    acc = mm(acc, zm1M);
    sc inv = sc_invert_mont_vartime(acc);   // public challenges: binary Euclid instead of x^(l-2)
    const sc zm1_inv = mm(inv, pre_zm1);           // Montgomery form
    inv = mm(inv, zm1M);
    const sc ym1_inv = mm(inv, pre_ym1);
    inv = mm(inv, ym1M);
    const sc yinv = sc_from_mont(mm(inv, pre_y));
    inv = mm(inv, yM);
    sc *dl = d + 4, *dr = d + 4 + lg;
#pragma unroll 1
    for (uint32_t j = lg; j-- > 0;) {
        sc uj = c[CH_UJ0 + j];
        sc ujinv = sc_from_mont(mm(inv, c[CH_UJ0 + lg + j]));
        inv = mm(inv, sc_to_mont(uj));
        sc w_l = sc_mul(uj, uj), w_r = sc_mul(ujinv, ujinv);
        __syncwarp();
        if (lead) { c[CH_UJ0 + lg + j] = ujinv; dl[j] = w_l; dr[j] = w_r; }
    }
    // geometric series: 1 + b + .. + b^(cnt-1) = (b^cnt - 1) / (b - 1) for cnt >= 64 and b != 1, summed directly otherwise
    const sc zM = sc_to_mont(z);
    sc sum_y, sum_z;
    if (nm >= 64 && !y_is_1) {
        sc e = yM;
        for (uint64_t k = 1; k < nm; k <<= 1) e = mm(e, e);      // n m is a power of two: y^(n m) by squarings
        sum_y = sc_from_mont(mm(sc_sub(e, sc_to_mont(one)), ym1_inv));
    } else {
        sc s = sc_zero(), e = sc_to_mont(one);
        for (uint64_t k = 0; k < nm; k++) { s = sc_add(s, e); e = mm(e, yM); }
        sum_y = sc_from_mont(s);
    }
    {
        sc s = sc_zero(), e = sc_to_mont(one);
        for (uint32_t k = 0; k < m; k++) { s = sc_add(s, e); e = mm(e, zM); }   // e = z^m afterwards
        sum_z = (m >= 64 && !z_is_1) ? sc_from_mont(mm(sc_sub(e, sc_to_mont(one)), zm1_inv)) : sc_from_mont(s);
    }
    sc sum_2 = sc_zero();
    if (nbits >= 64) { sum_2.v[0] = 0xffffffffu; sum_2.v[1] = 0xffffffffu; }   // 2^64 - 1 (the protocol's widest range)
    else { const uint64_t v = (1ull << nbits) - 1; sum_2.v[0] = (uint32_t)v; sum_2.v[1] = (uint32_t)(v >> 32); }
    const sc zz = sc_mul(z, z);
    const sc delta = sc_sub(sc_mul(sc_sub(z, zz), sum_y), sc_mul(sc_mul(sc_mul(zz, z), sum_2), sum_z));
    sc t_x, t_x_bl, e_bl, a, b;
    const uint32_t *sw = (const uint32_t *)(pf + 128);
    for (int k = 0; k < 8; k++) { t_x.v[k] = sw[k]; t_x_bl.v[k] = sw[8 + k]; e_bl.v[k] = sw[16 + k]; }
    const uint32_t *abw = (const uint32_t *)(LR + 64 * (size_t)lg);
    for (int k = 0; k < 8; k++) { a.v[k] = abw[k]; b.v[k] = abw[8 + k]; }
    const sc tx_coef = sc_add(sc_mul(w, sc_sub(t_x, sc_mul(a, b))), sc_mul(cc, sc_sub(delta, t_x)));
    const sc txbl_coef = sc_sub(sc_neg(e_bl), sc_mul(cc, t_x_bl));
    // commitment weights c z^(2+j)
    sc ez = sc_mul(cc, zz);
    for (uint32_t j = 0; j < m; j++) {
        if (lead) d[4 + 2 * lg + j] = ez;
        ez = sc_mul(ez, z);
    }
    if (!lead) return;
    c[CH_Y] = y; c[CH_YINV] = yinv; c[CH_Z] = z; c[CH_X] = x; c[CH_W] = w; c[CH_A] = a; c[CH_B] = b; c[CH_R] = rho; c[CH_RHO] = one;
    c[CH_TX] = tx_coef; c[CH_TXBL] = txbl_coef;
    c[CH_U] = sc_zero(); c[CH_UJ] = sc_zero(); c[CH_UJINV] = sc_zero(); c[CH_EBL] = sc_zero();
    d[0] = one; d[1] = x; d[2] = sc_mul(cc, x); d[3] = sc_mul(cc, sc_mul(x, x));
}

// per-pass weights of a range-proof verification: rho_k (combined pass) or 1 (per-request pass), 0 for a request with a point
// that did not decompress; applied to the B / B_blinding coefficients and the dynamic scalars, CH_RHO for k_rp_verify_scalars
__global__ void __launch_bounds__(128) k_rp_apply_weights(const sc *__restrict__ chal0, const sc *__restrict__ dyn0, const uint8_t *__restrict__ valid,
                                                          uint32_t ds, uint32_t n_req, uint32_t combined, sc *__restrict__ chal, sc *__restrict__ dyn) {
    const uint32_t p = blockIdx.x, t = threadIdx.x;
    __shared__ uint32_t dead;
    if (t == 0) dead = 0;
    __syncthreads();
    for (uint32_t k = t; k < ds; k += blockDim.x)
        if (!valid[(size_t)p * ds + k]) dead = 1;
    __syncthreads();
    const sc *c0 = chal0 + (size_t)p * CH_N;
    const sc rho = dead ? sc_zero() : (combined ? c0[CH_R] : sc_one());
    sc *c = chal + (size_t)p * CH_N;
    for (uint32_t k = t; k < CH_N; k += blockDim.x) {
        sc v = c0[k];
        if (k == CH_RHO) v = rho;
        else if (k == CH_TX || k == CH_TXBL) v = sc_mul(rho, v);
        c[k] = v;
    }
    for (uint32_t k = t; k < ds; k += blockDim.x) dyn[(size_t)p * ds + k] = sc_mul(rho, dyn0[(size_t)p * ds + k]);
}

// ---------------------------------------------------------------- batch-verification weights on the device
// The random linear combination of a batch verification needs one secret, proof-binding weight per request. Deriving them
// on the host meant a device -> host -> device round trip in the middle of every batch (the digests r_i come out of the
// transcript replay, the weights go into the scalar assembly) plus ~B serial Keccak permutations on one host thread. Here
// they are a two-level SHAKE256 tree, all on the launching stream (hashlib-reproducible: tests/test_gpu_protocol.py):
//   h_g   = SHAKE256(LBL_H || LE64(g) || LE64(cnt_g) || r_{32g} || ... || r_{32g+cnt_g-1})[0..32)        one warp per 32 requests
//   D     = SHAKE256(LBL_D || batch_seed || LE64(B) || h_0 || ... || h_{G-1})[0..32)                      one warp (per block)
//   rho_i = from_bytes_mod_order_wide(SHAKE256(LBL_W || D || LE64(i))[0..64))                              one thread per request
// r_i = the verifier scalar request i's own transcript yields after absorbing its whole proof (binds every proof byte),
// batch_seed = the caller's secret. A request with a point that failed to decompress gets weight zero (its status is
// reported separately). Labels are 32 bytes, zero padded.
__device__ __forceinline__ uint64_t bw_label_word(int which, uint32_t w) {
    // "bbp batch digest v1", "bbp batch root v1", "bbp batch weight v1" as little-endian 64-bit words
    const char *s = which == 0 ? "bbp batch digest v1\0\0\0\0\0\0\0\0\0\0\0\0\0" : which == 1 ? "bbp batch root v1\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0" : "bbp batch weight v1\0\0\0\0\0\0\0\0\0\0\0\0\0";
    uint64_t v = 0;
    for (int k = 7; k >= 0; k--) v = (v << 8) | (uint8_t)s[8 * w + k];
    return v;
}
// SHAKE256 (rate 136 = 17 lanes) of a message given as 64-bit words (n_words of them), state spread over the warp's lanes
// (keccak_warp); returns this lane's state word after the last permutation (lanes 0..7 = the first 64 output bytes).
template <class F>
__device__ __forceinline__ uint64_t shake256_warp_words(const keccak_warp &K, uint32_t lane, uint32_t n_words, F word) {
    uint64_t a = 0;
    const uint32_t n_blocks = n_words / 17 + 1;
#pragma unroll 1
    for (uint32_t blk = 0; blk < n_blocks; blk++) {
        uint64_t w = 0;
        if (lane < 17) {
            const uint32_t wi = 17 * blk + lane;
            if (wi < n_words) w = word(wi);
            else if (wi == n_words) w = 0x1full;                           // SHAKE domain bits + first pad bit
            if (blk == n_blocks - 1 && lane == 16) w ^= 0x80ull << 56;     // last pad bit
        }
        a = K.permute(a ^ w);
    }
    return a;
}
__global__ void __launch_bounds__(32) k_batch_weight_digests(const sc *__restrict__ chal, uint32_t n_req, uint64_t *__restrict__ digests) {
    const uint32_t g = blockIdx.x, lane = threadIdx.x;
    const uint32_t cnt = min(32u, n_req - 32 * g);
    keccak_warp K;
    K.init(lane);
    const uint64_t a = shake256_warp_words(K, lane, 6 + 4 * cnt, [&](uint32_t w) -> uint64_t {
        if (w < 4) return bw_label_word(0, w);
        if (w == 4) return g;
        if (w == 5) return cnt;
        const uint32_t i = 32 * g + (w - 6) / 4, k = (w - 6) % 4;
        const sc &r = chal[(size_t)i * CH_N + CH_R];
        return (uint64_t)r.v[2 * k] | ((uint64_t)r.v[2 * k + 1] << 32);
    });
    if (lane < 4) digests[4 * (size_t)g + lane] = a;
}
// combined = 0: weights 1 (0 for a request with an invalid point), no hashing
__global__ void __launch_bounds__(128) k_batch_weights(sc *__restrict__ chal, const uint8_t *__restrict__ valid, uint32_t ds, uint32_t n_req,
                                                       const uint64_t *__restrict__ digests, const uint8_t *__restrict__ seed, uint32_t combined) {
    __shared__ uint64_t root[4];
    const uint32_t lane = threadIdx.x & 31;
    if (combined && threadIdx.x < 32) {
        keccak_warp K;
        K.init(lane);
        const uint32_t G = (n_req + 31) / 32;
        const uint64_t *s64 = (const uint64_t *)seed;
        const uint64_t a = shake256_warp_words(K, lane, 9 + 4 * G, [&](uint32_t w) -> uint64_t {
            if (w < 4) return bw_label_word(1, w);
            if (w < 8) return s64[w - 4];
            if (w == 8) return n_req;
            return digests[w - 9];
        });
        if (lane < 4) root[lane] = a;
    }
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_req) return;
    bool alive = true;
    for (uint32_t k = 0; k < ds; k++) alive = alive && valid[(size_t)i * ds + k] != 0;
    sc rho = sc_one();
    if (combined) {
        uint64_t st[25];
#pragma unroll
        for (int k = 0; k < 25; k++) st[k] = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { st[k] = bw_label_word(2, k); st[4 + k] = root[k]; }
        st[8] = i;
        st[9] = 0x1full;
        st[16] = 0x80ull << 56;
        keccak_f1600_dev(st);
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 8; k++) { w[2 * k] = (uint32_t)st[k]; w[2 * k + 1] = (uint32_t)(st[k] >> 32); }
        rho = sc_from_wide_words(w);
    }
    chal[(size_t)i * CH_N + CH_RHO] = alive ? rho : sc_zero();
}

}  // namespace bbp
