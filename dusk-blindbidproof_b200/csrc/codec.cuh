// Batch point codecs: Ristretto decompress / compress, hash-to-group, and conversions into the 96-byte affine
// niels layout the MSM kernels stream (Appendix A). One thread per point; each does one field
// exponentiation (sqrt_ratio_i or inversion), so these kernels are pure FMA-pipe work.
#pragma once
#include "ge25519.cuh"

namespace bbp {

// compressed (n x 32 B) -> niels (n x 96 B); invalid encodings clear *all_valid and write the identity
__global__ void __launch_bounds__(128) k_decompress_to_niels(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, uint32_t n,
                                                             int *__restrict__ all_valid, uint8_t *__restrict__ valid_flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[8];
    const uint4 *q = (const uint4 *)(in + 8 * (size_t)i);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    ge p;
    bool ok = ge_decompress_words(p, w);
    if (!ok) { p = ge_identity(); atomicAnd(all_valid, 0); }
    if (valid_flags) valid_flags[i] = ok ? 1 : 0;
    niels_store(out + 96 * (size_t)i, ge_affine_to_niels(p.X, p.Y, p.T));
}

// same, reading point i from a strided layout: groups of `per_group` consecutive 32-byte encodings, one group every
// `group_stride` bytes (the verifier's request blobs)
__global__ void __launch_bounds__(128) k_decompress_to_niels_strided(const uint8_t *__restrict__ in, uint32_t per_group, uint32_t group_stride,
                                                                     uint8_t *__restrict__ out, uint32_t n, uint8_t *__restrict__ valid_flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *q = (const uint4 *)(in + (size_t)(i / per_group) * group_stride + 32 * (size_t)(i % per_group));
    uint4 a = __ldg(q), b = __ldg(q + 1);
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    ge p;
    bool ok = ge_decompress_words(p, w);
    if (!ok) p = ge_identity();
    valid_flags[i] = ok ? 1 : 0;
    niels_store(out + 96 * (size_t)i, ge_affine_to_niels(p.X, p.Y, p.T));
}

// compressed -> extended (n x 128 B)
__global__ void __launch_bounds__(128) k_decompress_to_ext(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, uint32_t n,
                                                           int *__restrict__ all_valid, uint8_t *__restrict__ valid_flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[8];
    const uint4 *q = (const uint4 *)(in + 8 * (size_t)i);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    ge p;
    bool ok = ge_decompress_words(p, w);
    if (!ok) { p = ge_identity(); atomicAnd(all_valid, 0); }
    if (valid_flags) valid_flags[i] = ok ? 1 : 0;
    // API-facing: canonical field encodings (internal tables keep the lazy representation)
    p.X = fe_canon(p.X); p.Y = fe_canon(p.Y); p.Z = fe_canon(p.Z); p.T = fe_canon(p.T);
    ge_store(out + 128 * (size_t)i, p);
}

// extended (n x 128 B, any Z != 0) -> niels. Montgomery's trick over K points per thread: one inversion per K.
#define BBP_NIELS_BATCH 8
__global__ void __launch_bounds__(128) k_ext_to_niels(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint32_t n) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i0 = t * BBP_NIELS_BATCH;
    if (i0 >= n) return;
    uint32_t cnt = min((uint32_t)BBP_NIELS_BATCH, n - i0);
    fe prefix[BBP_NIELS_BATCH];
    fe acc = fe_one();
#pragma unroll 1
    for (uint32_t k = 0; k < cnt; k++) {
        prefix[k] = acc;
        acc = fe_mul(acc, fe_load(in + 128 * (size_t)(i0 + k) + 64));
    }
    fe inv = fe_invert(acc);
#pragma unroll 1
    for (uint32_t k = cnt; k-- > 0;) {
        ge p = ge_load(in + 128 * (size_t)(i0 + k));
        fe zinv = fe_mul(inv, prefix[k]);
        inv = fe_mul(inv, p.Z);
        niels_store(out + 96 * (size_t)(i0 + k), ge_to_niels(p, zinv));
    }
}

// extended -> compressed
__global__ void __launch_bounds__(128) k_compress(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ge p = ge_load(in + 128 * (size_t)i);
    uint32_t w[8];
    ge_compress_words(w, p);
    uint4 *q = (uint4 *)(out + 8 * (size_t)i);
    q[0] = make_uint4(w[0], w[1], w[2], w[3]);
    q[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// 64 uniform bytes -> extended point (RistrettoPoint::from_uniform_bytes)
__global__ void __launch_bounds__(128) k_from_uniform(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = in[16 * (size_t)i + k];
    ge_store(out + 128 * (size_t)i, ge_from_uniform_words(w));
}

// fixed-base window table: row w of `out` (stride entries of 96 B) = niels(2^(c*w) * P_i), built from extended inputs.
// One thread per point walks the windows (c doublings each) and normalises each multiple with its own inversion.
__global__ void __launch_bounds__(128) k_build_window_table(const uint8_t *__restrict__ in_ext, uint8_t *__restrict__ out, uint32_t n, uint32_t c,
                                                            uint32_t W, uint32_t stride) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ge p = ge_load(in_ext + 128 * (size_t)i);
#pragma unroll 1
    for (uint32_t w = 0; w < W; w++) {
        fe zinv = fe_invert(p.Z);
        niels_store(out + 96 * ((size_t)w * stride + i), ge_to_niels(p, zinv));
        if (w + 1 < W) {
#pragma unroll 1
            for (uint32_t k = 0; k < c; k++) p = ge_dbl(p);
        }
    }
}

// integer-pipe ceiling: BBP_PEAK_ILP independent chains of the IMAD.WIDE.U32 (32x32+64) the field multiplier is built
// from (ptxas fuses each mad.lo.cc / madc.hi pair into one), at full occupancy. The roofline denominator of the MSM.
#define BBP_PEAK_ILP 8
#define BBP_PEAK_ITERS 4096
__global__ void __launch_bounds__(1024, 2) k_int_peak(uint32_t *out, uint32_t seed, unsigned long long *cycles) {
    uint64_t acc[BBP_PEAK_ILP];
    uint32_t x = seed + threadIdx.x;
#pragma unroll
    for (int i = 0; i < BBP_PEAK_ILP; i++) acc[i] = ((uint64_t)(seed * 3 + blockIdx.x + 7 * i) << 32) | (x + i);
    __syncthreads();
    long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < BBP_PEAK_ITERS; it += 4) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < BBP_PEAK_ILP; i++) acc[i] += (uint64_t)(uint32_t)acc[i] * x;   // IMAD.WIDE.U32 Rd, Ra.lo, x, Rd
        }
    }
    long long c1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < BBP_PEAK_ILP; i++) r ^= (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(c1 - c0);
}

// the same measurement on the instruction pair the field multiplier is actually written in: mad.lo.cc.u32 / madc.hi.u32 on
// a 64-bit accumulator (ptxas fuses the pair into one IMAD.WIDE.U32; if it ever did not, this variant would show it)
__global__ void __launch_bounds__(1024, 2) k_int_peak_pair(uint32_t *out, uint32_t seed, unsigned long long *cycles) {
    uint32_t lo[BBP_PEAK_ILP], hi[BBP_PEAK_ILP];
    uint32_t x = seed + threadIdx.x;
#pragma unroll
    for (int i = 0; i < BBP_PEAK_ILP; i++) { lo[i] = x + i; hi[i] = seed * 3 + blockIdx.x + 7 * i; }
    __syncthreads();
    long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < BBP_PEAK_ITERS; it += 4) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < BBP_PEAK_ILP; i++)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(lo[(i + 1) % BBP_PEAK_ILP] | 1u), "r"(x));
        }
    }
    long long c1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < BBP_PEAK_ILP; i++) r ^= lo[i] ^ hi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(c1 - c0);
}

// sharded IPP: slot s of every rank's partial sums (rank-major, n_slots x 128 B per rank) -> one compressed point per slot
__global__ void __launch_bounds__(64) k_sum_ranks_compress(const uint8_t *__restrict__ in, uint32_t n_slots, uint32_t world, uint32_t *__restrict__ out) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    ge acc = ge_load(in + 128 * (size_t)s);
    for (uint32_t r = 1; r < world; r++) acc = ge_add(acc, ge_load(in + 128 * ((size_t)r * n_slots + s)));
    ge_compress_words(out + 8 * (size_t)s, acc);
}

// sum of n (<= 1024) extended points, compressed: the local tail of a sharded MSM after the all-gather of the
// per-GPU partial sums. One warp: lane sums, then a shuffle-free tree through shared memory.
__global__ void __launch_bounds__(32) k_sum_compress(const uint8_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ out) {
    __shared__ uint4 sm_u4[32 * 8];
    uint8_t *sm = (uint8_t *)sm_u4;
    uint32_t lane = threadIdx.x;
    ge acc = ge_identity();
    for (uint32_t i = lane; i < n; i += 32) acc = ge_add(acc, ge_load(in + 128 * (size_t)i));
    ge_store(sm + 128 * lane, acc);
    __syncwarp();
    for (uint32_t stride = 16; stride >= 1; stride >>= 1) {
        if (lane < stride) ge_store(sm + 128 * lane, ge_add(ge_load(sm + 128 * lane), ge_load(sm + 128 * (lane + stride))));
        __syncwarp();
    }
    if (lane == 0) ge_compress_words(out, ge_load(sm));
}

__global__ void k_store_basepoint(uint8_t *out) { ge_store(out, ge_basepoint()); }

// ---- unit-test kernels (driven by the bbp_test_* entry points; compared against the oracle in tests/)
// op: 0 mul, 1 add, 2 sub, 3 invert(a), 4 sq(a), 5 neg(a)
__global__ void k_test_fe(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint32_t *__restrict__ out, uint32_t n, int op) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y, r;
#pragma unroll
    for (int k = 0; k < 8; k++) { x.v[k] = a[8 * (size_t)i + k]; y.v[k] = b[8 * (size_t)i + k]; }
    switch (op) {
        case 0: r = fe_mul(x, y); break;
        case 1: r = fe_add(x, y); break;
        case 2: r = fe_sub(x, y); break;
        case 3: r = fe_invert(x); break;
        case 4: r = fe_sq(x); break;
        default: r = fe_neg(x); break;
    }
    fe_tobytes_words(out + 8 * (size_t)i, r);
}
// op: 0 add (ext+ext), 1 double, 2 madd via niels of b, 3 msub via niels of b, 4 from_niels(b)
__global__ void k_test_ge(const uint8_t *__restrict__ a_ext, const uint8_t *__restrict__ b_ext, uint8_t *__restrict__ out_ext, uint32_t n, int op) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ge p = ge_load(a_ext + 128 * (size_t)i), q = ge_load(b_ext + 128 * (size_t)i), r;
    niels qn = ge_to_niels(q, fe_invert(q.Z));
    switch (op) {
        case 0: r = ge_add(p, q); break;
        case 1: r = ge_dbl(p); break;
        case 2: r = ge_madd(p, qn, false); break;
        case 3: r = ge_madd(p, qn, true); break;
        default: r = ge_from_niels(qn, false); break;
    }
    ge_store(out_ext + 128 * (size_t)i, r);
}

}  // namespace bbp
