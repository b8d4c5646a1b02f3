// Latency path for small fixed-base MSMs over the context's generators (single requests and batches of a few).
//
// The bucket engine (msm.cuh) is built for throughput: ~20 launches per MSM (recode, scans, counting sort, task table,
// accumulation, two reduction levels, Horner), each of which costs 10-30 us however little work it has, so the 11
// inner-product rounds of a single proof spent 0.5 ms each moving 2 x 2049 terms. Here the generators carry a second
// table with the digit multiples themselves — for every generator G and 5-bit window w the sixteen points
// d * 2^(5w) * G, d = 1..16, in affine niels form (51 windows x 16 x 96 B = 78 KB per generator, 321 MB for the 4098
// generators of BulletproofGens::new(2048, 1)) — so that a term is 51 table look-ups and mixed additions with no buckets,
// no sort and no doublings: one kernel adds (term, window group) pairs and folds each block to one partial sum, a second
// one folds the partials of a slot and compresses. Same group element as the engine, hence the same bytes.
// Replaces, for small batches, what the reference computes with dalek's Straus / Pippenger
// (the MSMs under Prover::prove / Verifier::verify, src/blindbid/proof.rs:88,
// src/blindbid/verify.rs:88).
#pragma once
#include "ge25519.cuh"
#include "sc25519.cuh"

namespace bbp {

static const uint32_t SM_C = 5, SM_W = 51, SM_D = 16;    // window bits, windows (255 >= 253 bits), multiples per window
static const uint32_t SM_GROUPS = 4;                     // warps per block; warp g takes the windows w = g (mod 4) of the block's 32 terms
static const uint32_t SM_THREADS = 32 * SM_GROUPS;

// out[((w * n + i) * 16 + (d - 1))] = niels(d * 2^(5w) * P_i); one thread per (w, i)
__global__ void __launch_bounds__(128) k_build_digit_table(const uint8_t *__restrict__ in_ext, uint8_t *__restrict__ out, uint32_t n) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * SM_W) return;
    uint32_t w = idx / n, i = idx % n;
    ge base = ge_load(in_ext + 128 * (size_t)i);
#pragma unroll 1
    for (uint32_t k = 0; k < SM_C * w; k++) base = ge_dbl(base);
    ge m = base;
    uint8_t *o = out + 96 * (size_t)idx * SM_D;
#pragma unroll 1
    for (uint32_t d = 1; d <= SM_D; d++) {
        niels_store(o + 96 * (size_t)(d - 1), ge_to_niels(m, fe_invert(m.Z)));
        if (d < SM_D) m = ge_add(m, base);
    }
}

// 5 bits of a 256-bit little-endian integer at bit offset b (b + 5 <= 256)
__device__ __forceinline__ uint32_t sm_bits5(const uint32_t *v, uint32_t b) {
    uint32_t limb = b >> 5, sh = b & 31;
    uint32_t lo = v[limb], hi = limb < 7 ? v[limb + 1] : 0u;
    return __funnelshift_r(lo, hi, sh) & 31u;
}

__device__ inline ge sm_block_fold(ge acc, ge *sh) {
    const uint32_t t = threadIdx.x;
    sh[t] = acc;
    __syncthreads();
#pragma unroll 1
    for (uint32_t s = SM_THREADS / 2; s >= 1; s >>= 1) {
        if (t < s) sh[t] = ge_add(sh[t], sh[t + s]);
        __syncthreads();
    }
    return sh[0];
}

// grid (chunks, slots); block = 4 window groups (warps) x 32 terms. scalars: [n_slots][slot_len] reduced, normal form.
// Entry e of a slot multiplies generator column colmap[(slot % colmap_slots) * slot_len + e] (or e without a map).
__global__ void __launch_bounds__(SM_THREADS) k_small_msm_partial(const sc *__restrict__ scalars, uint32_t slot_len, const uint8_t *__restrict__ table,
                                                                   uint32_t n_gens, const uint32_t *__restrict__ colmap, uint32_t colmap_slots,
                                                                   uint8_t *__restrict__ partial, uint32_t uniform) {
    __shared__ ge sh[SM_THREADS];
    // the window group is the warp index, so that a warp's lanes (32 different terms) all add at the same windows
    const uint32_t slot = blockIdx.y, t = threadIdx.x, g = t >> 5;
    const uint32_t e = blockIdx.x * (SM_THREADS / SM_GROUPS) + (t & 31);
    ge acc = ge_identity();
    if (e < slot_len) {
        sc s = scalars[(size_t)slot * slot_len + e];
        if (s.v[7] >> 29) s = sc_reduce_words(s.v);   // callers of the raw MSM surface may pass any 256-bit value; the recoding needs < 2^253
        if (uniform) {
            // secret scalars (BBP_CT_COMMIT): the same instruction stream and the same number of table reads for every
            // scalar value — no zero-scalar exit, no zero-digit skip; a zero digit reads entry 16 and its sum is discarded
            // by a limb-wise select. (The table ADDRESS still depends on the digit; see DESIGN.md §constant time.)
            const uint32_t col = colmap ? colmap[(size_t)(slot % colmap_slots) * slot_len + e] : e;
            uint32_t carry = 0;
#pragma unroll 1
            for (uint32_t w = 0; w < SM_W; w++) {
                uint32_t v = sm_bits5(s.v, SM_C * w) + carry;
                carry = v > SM_D ? 1u : 0u;
                int d = (int)v - (int)(carry << SM_C);
                if ((w & (SM_GROUPS - 1)) == g) {      // public: depends on the window index only
                    uint32_t a = (uint32_t)(d < 0 ? -d : d);
                    const uint32_t keep = 0u - (uint32_t)(d != 0);
                    niels q = niels_load_ro(table + 96 * (((size_t)w * n_gens + col) * SM_D + ((a - 1) & (SM_D - 1))));
                    ge sum = ge_madd(acc, q, d < 0);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        acc.X.v[k] = (sum.X.v[k] & keep) | (acc.X.v[k] & ~keep); acc.Y.v[k] = (sum.Y.v[k] & keep) | (acc.Y.v[k] & ~keep);
                        acc.Z.v[k] = (sum.Z.v[k] & keep) | (acc.Z.v[k] & ~keep); acc.T.v[k] = (sum.T.v[k] & keep) | (acc.T.v[k] & ~keep);
                    }
                }
            }
        } else if (!sc_iszero(s)) {
            const uint32_t col = colmap ? colmap[(size_t)(slot % colmap_slots) * slot_len + e] : e;
            uint32_t carry = 0;
#pragma unroll 1
            for (uint32_t w = 0; w < SM_W; w++) {
                uint32_t v = sm_bits5(s.v, SM_C * w) + carry;
                carry = v > SM_D ? 1u : 0u;
                int d = (int)v - (int)(carry << SM_C);
                if ((w & (SM_GROUPS - 1)) == g && d != 0) {
                    uint32_t a = (uint32_t)(d < 0 ? -d : d);
                    niels q = niels_load_ro(table + 96 * (((size_t)w * n_gens + col) * SM_D + (a - 1)));
                    acc = ge_madd(acc, q, d < 0);
                }
            }
        }
    }
    acc = sm_block_fold(acc, sh);
    if (t == 0) ge_store(partial + 128 * ((size_t)slot * gridDim.x + blockIdx.x), acc);
}

// one block per slot: folds the slot's partial sums; writes the extended point and / or the compressed bytes
__global__ void __launch_bounds__(SM_THREADS) k_small_msm_final(const uint8_t *__restrict__ partial, uint32_t n_partial, uint8_t *__restrict__ out_ext,
                                                                 uint8_t *__restrict__ out_compressed) {
    __shared__ ge sh[SM_THREADS];
    const uint32_t slot = blockIdx.x, t = threadIdx.x;
    ge acc = ge_identity();
    for (uint32_t k = t; k < n_partial; k += SM_THREADS) acc = ge_add(acc, ge_load(partial + 128 * ((size_t)slot * n_partial + k)));
    acc = sm_block_fold(acc, sh);
    if (t == 0) {
        if (out_ext) ge_store(out_ext + 128 * (size_t)slot, acc);
        if (out_compressed) ge_compress_words((uint32_t *)(out_compressed + 32 * (size_t)slot), acc);
    }
}

}  // namespace bbp
