// Fixed-base machinery over the resident generator set.
//  * comb table for the two Pedersen bases (B, B_blinding): every Prover::commit / T_k commitment of the reference is
//    v*B + r*B_blinding (bulletproofs PedersenGens::commit, called from src/blindbid/proof.rs:55-67 and inside
//    Prover::prove). With 64 radix-16 windows x 8 multiples per base a commitment is
//    128 table lookups + mixed additions and no doubling; one thread per commitment, compressed in the same thread.
//  * window table over ALL generators (2^(c*w) * P_i, affine niels): lets the Pippenger engine put every window of a
//    slot into ONE bucket set (msm.cuh "fixed" layout), so the protocol's MSMs over G / H never double.
#pragma once
#include "ge25519.cuh"
#include "sc25519.cuh"

namespace bbp {

#define BBP_COMB_WINDOWS 64
#define BBP_COMB_ENTRIES (2 * BBP_COMB_WINDOWS * 8)

// thread t = base * 64 + w builds k * 16^w * P_base for k = 1..8
__global__ void __launch_bounds__(128) k_build_comb(const uint8_t *__restrict__ gens_ext, uint8_t *__restrict__ table) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * BBP_COMB_WINDOWS) return;
    uint32_t base = t / BBP_COMB_WINDOWS, w = t % BBP_COMB_WINDOWS;
    ge p = ge_load(gens_ext + 128 * (size_t)base);
#pragma unroll 1
    for (uint32_t k = 0; k < 4 * w; k++) p = ge_dbl(p);
    ge m = p;
#pragma unroll 1
    for (uint32_t k = 0; k < 8; k++) {
        niels_store(table + 96 * ((size_t)t * 8 + k), ge_to_niels(m, fe_invert(m.Z)));
        m = ge_add(m, p);
    }
}

// vals: n x 2 reduced scalars (value, blinding); out: n x 32 B compressed v*B + r*B_blinding
__global__ void __launch_bounds__(128) k_pedersen_commit(const sc *__restrict__ vals, const uint8_t *__restrict__ comb, uint32_t *__restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ge acc = ge_identity();
#pragma unroll 1
    for (uint32_t base = 0; base < 2; base++) {
        sc s = vals[2 * (size_t)i + base];
        uint32_t carry = 0;
#pragma unroll 1
        for (uint32_t w = 0; w < BBP_COMB_WINDOWS; w++) {
            uint32_t nib = ((s.v[w >> 3] >> (4 * (w & 7))) & 15u) + carry;
            carry = nib > 8 ? 1u : 0u;
            int32_t d = (int32_t)nib - (int32_t)(carry << 4);
            if (d != 0) {
                uint32_t mag = (uint32_t)(d < 0 ? -d : d);
                niels q = niels_load_ro(comb + 96 * ((size_t)(base * BBP_COMB_WINDOWS + w) * 8 + (mag - 1)));
                acc = ge_madd(acc, q, d < 0);
            }
        }
    }
    ge_compress_words(out + 8 * (size_t)i, acc);
}

// sum of `count` extended points per group (groups are contiguous), identity test and optional compression:
// flags[g] = 1 if the sum is the Ristretto identity. One thread per group.
__global__ void __launch_bounds__(64) k_group_sum_identity(const uint8_t *__restrict__ pts_ext, uint32_t n_groups, uint32_t count, uint32_t group_stride,
                                                            uint8_t *__restrict__ flags, uint32_t *__restrict__ out_compressed) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    ge acc = ge_load(pts_ext + 128 * (size_t)g);
    for (uint32_t k = 1; k < count; k++) acc = ge_add(acc, ge_load(pts_ext + 128 * ((size_t)k * group_stride + g)));
    if (flags) flags[g] = ge_is_identity(acc) ? 1 : 0;
    if (out_compressed) ge_compress_words(out_compressed + 8 * (size_t)g, acc);
}

// out[g] = sum over parts of partial[part][g], g < n_groups (extended points, 128 B each): folds the per-lane partial
// sums of a sharded batch verification into the one this GPU contributes
__global__ void __launch_bounds__(32) k_partial_fold(const uint8_t *__restrict__ partial, uint32_t n_parts, uint32_t n_groups, uint8_t *__restrict__ out) {
    uint32_t g = threadIdx.x;
    if (g >= n_groups) return;
    ge acc = ge_load(partial + 128 * (size_t)g);
    for (uint32_t k = 1; k < n_parts; k++) acc = ge_add(acc, ge_load(partial + 128 * ((size_t)k * n_groups + g)));
    ge_store(out + 128 * (size_t)g, acc);
}

// Verdict of a proof-range sharded batch verification after the all-gather: rows[r] = rank r's two extended partial sums
// (2 x 128 B) followed by its local flag byte. out[0] = 1 iff every flag is set and the sum of the 2 x world points is the
// Ristretto identity — tested in extended coordinates, no compression (its field exponentiation was most of the old
// sum-and-compress tail). One warp: lane sums, then a tree through shared memory.
__global__ void __launch_bounds__(32) k_sharded_verdict(const uint8_t *__restrict__ rows, uint32_t world, uint32_t row_stride, uint8_t *__restrict__ out) {
    __shared__ uint4 sm_u4[32 * 8];
    uint8_t *sm = (uint8_t *)sm_u4;
    const uint32_t lane = threadIdx.x;
    ge acc = ge_identity();
    bool ok = true;
    for (uint32_t i = lane; i < 2 * world; i += 32) acc = ge_add(acc, ge_load(rows + (size_t)(i >> 1) * row_stride + 128 * (i & 1)));
    for (uint32_t r = lane; r < world; r += 32) ok = ok && rows[(size_t)r * row_stride + 256] != 0;
    ge_store(sm + 128 * lane, acc);
    __syncwarp();
    for (uint32_t stride = 16; stride >= 1; stride >>= 1) {
        if (lane < stride) ge_store(sm + 128 * lane, ge_add(ge_load(sm + 128 * lane), ge_load(sm + 128 * (lane + stride))));
        __syncwarp();
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) out[0] = (all_ok && ge_is_identity(ge_load(sm))) ? 1 : 0;
}

}  // namespace bbp
