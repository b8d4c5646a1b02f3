// Signed-digit Pippenger multiscalar multiplication on one B200: sum_i s_i * P_i over Ristretto255.
// Replaces curve25519-dalek 1.2.3's VartimeMultiscalarMul / MultiscalarMul as bulletproofs calls them underneath
// src/blindbid/proof.rs:88 and src/blindbid/verify.rs:88. The compressed
// result is algorithm independent, so the decomposition below is free to be GPU shaped.
//
// Pipeline (all on one stream, no host synchronisation between stages):
//   1 recode     scalar -> mod l -> W signed radix-2^c digits; histogram of (slot, window, |digit|) keys
//   2 scan       exclusive scan of the histogram (bucket offsets) and of ceil(count / S) (task offsets)
//   3 scatter    counting sort of (point ref | sign) entries by key
//   4 accumulate one thread per task (<= S entries of one bucket): 7M mixed additions from the niels table
//   5 reduce     heavy buckets folded by a warp each; then levels of 8-way merges of (sum, index-weighted sum) pairs
//   6 combine    Horner over windows (c doublings each, four lanes per slot), Ristretto compression
// Two base layouts share the kernels: "variable" bases (one niels entry per point, W bucket sets per slot) and
// "fixed" bases (a precomputed table of 2^(c*w) * P_i, all windows of a slot share ONE bucket set, no doublings).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ge25519.cuh"
#include "sc25519.cuh"

namespace bbp {

#define BBP_CUDA_OK(expr)                                                                  \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            fprintf(stderr, "bbp: CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return -100;                                                                   \
        }                                                                                  \
    } while (0)

struct msm_shape {
    uint32_t n;            // scalars in this launch (all slots together)
    uint32_t n_per_slot;   // scalars per slot (slot = i / n_per_slot)
    uint32_t n_slots;
    uint32_t base_mod;     // point ref of scalar i = i % base_mod (bases shared between slots) ; = n for distinct bases
    uint32_t c;            // window bits
    uint32_t W;            // windows
    uint32_t B;            // buckets per bucket set = 2^(c-1)
    uint32_t fixed;        // 1: table[w * table_stride + ref] holds 2^(c*w) * P_ref, one bucket set per slot
    uint32_t table_stride; // entries per window row of a fixed table
    uint32_t sets_per_slot;// W (variable) or 1 (fixed)
    uint32_t nkeys;        // n_slots * sets_per_slot * B
    uint32_t S;            // max entries per task
    // point reference of scalar i. mode 0: i % base_mod. mode 1: colmap[i % colmap_len] (compact slots over scattered
    // columns of a fixed table). mode 2: per-group tables — element e = i % n_per_slot of a slot refers to entry
    // (i / grp_div) * grp_stride + e, except the last element of every slot, which refers to the shared entry tail_ref.
    uint32_t ref_mode;
    const uint32_t *colmap;
    uint32_t colmap_len, grp_div, grp_stride, tail_ref;
};
__device__ __forceinline__ uint32_t msm_ref(const msm_shape &sh, uint32_t i) {
    if (sh.ref_mode == 1) return sh.colmap[i % sh.colmap_len];
    if (sh.ref_mode == 2) {
        uint32_t e = i % sh.n_per_slot;
        return (e + 1 == sh.n_per_slot) ? sh.tail_ref : (i / sh.grp_div) * sh.grp_stride + e;
    }
    return i % sh.base_mod;
}

// ---------------------------------------------------------------- recode + histogram
// digits are stored window-major: digits[w * n + i]
__global__ void k_recode(const uint32_t *__restrict__ scalars, int32_t *__restrict__ digits, uint32_t *__restrict__ hist, msm_shape sh) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sh.n) return;
    uint32_t w8[8];
    const uint4 *q = (const uint4 *)(scalars + 8 * (size_t)i);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    w8[0] = a.x; w8[1] = a.y; w8[2] = a.z; w8[3] = a.w; w8[4] = b.x; w8[5] = b.y; w8[6] = b.z; w8[7] = b.w;
    if ((w8[0] | w8[1] | w8[2] | w8[3] | w8[4] | w8[5] | w8[6] | w8[7]) == 0) {   // protocol slots are padded with zeros
        for (uint32_t w = 0; w < sh.W; w++) digits[(size_t)w * sh.n + i] = 0;
        return;
    }
    sc s = sc_reduce_words(w8);
    uint32_t slot = i / sh.n_per_slot;
    uint32_t carry = 0;
    const uint32_t c = sh.c, half = 1u << (c - 1), mask = (1u << c) - 1;
    for (uint32_t w = 0; w < sh.W; w++) {
        uint32_t bit = w * c, word = bit >> 5, off = bit & 31;
        uint32_t chunk = 0;
        if (word < 8) {
            chunk = s.v[word] >> off;
            if (off + c > 32 && word < 7) chunk |= s.v[word + 1] << (32 - off);
        }
        chunk = (chunk & mask) + carry;
        carry = (chunk > half) ? 1u : 0u;                   // digits in [-(2^(c-1) - 1), 2^(c-1)]
        int32_t d = (int32_t)chunk - (int32_t)(carry << c);
        digits[(size_t)w * sh.n + i] = d;
        if (d != 0) {
            uint32_t mag = (uint32_t)(d < 0 ? -d : d) - 1;
            uint32_t set = sh.fixed ? slot : slot * sh.W + w;
            atomicAdd(&hist[set * sh.B + mag], 1u);
        }
    }
}

// ---------------------------------------------------------------- exclusive scan (tiles of 4096)
// mode 0: in[i]; mode 1: ceil(in[i] / S)
__device__ __forceinline__ uint32_t scan_xform(uint32_t x, uint32_t S) { return S ? (x + S - 1) / S : x; }

__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t *__restrict__ tile_sums,
                                                     uint32_t n, uint32_t S) {
    __shared__ uint32_t warp_sums[32];
    uint32_t base = blockIdx.x * 4096u + threadIdx.x * 4u;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (base + k < n) ? scan_xform(in[base + k], S) : 0u;
    uint32_t tsum = v[0] + v[1] + v[2] + v[3];
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = warp_sums[lane], winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += t;
        }
        warp_sums[lane] = winc - ws;
        if (lane == 31) tile_sums[blockIdx.x] = winc;
    }
    __syncthreads();
    uint32_t ex = warp_sums[wid] + inc - tsum;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}
// single block: exclusive scan of the tile sums in place; writes the grand total to total_out
__global__ void __launch_bounds__(1024) k_scan_tile_sums(uint32_t *__restrict__ tile_sums, uint32_t n_tiles, uint32_t *__restrict__ total_out) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < n_tiles ? tile_sums[i] : 0u, inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane], winc = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= (uint32_t)o) winc += t;
            }
            warp_sums[lane] = winc - ws;
        }
        __syncthreads();
        uint32_t carry = carry_s;
        if (i < n_tiles) tile_sums[i] = carry + warp_sums[wid] + inc - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sums[wid] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}
__global__ void k_scan_add(uint32_t *__restrict__ out, const uint32_t *__restrict__ tile_sums, uint32_t n, const uint32_t *__restrict__ total) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_sums[i >> 12];
    if (i == n) out[n] = *total;   // one extra slot: out[n] = grand total
}

// ---------------------------------------------------------------- scatter (counting sort) + task table
__global__ void k_scatter(const int32_t *__restrict__ digits, const uint32_t *__restrict__ offs, uint32_t *__restrict__ cursor,
                          uint32_t *__restrict__ entries, msm_shape sh) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sh.n) return;
    uint32_t slot = i / sh.n_per_slot;
    uint32_t ref = msm_ref(sh, i);
    for (uint32_t w = 0; w < sh.W; w++) {
        int32_t d = digits[(size_t)w * sh.n + i];
        if (d == 0) continue;
        uint32_t mag = (uint32_t)(d < 0 ? -d : d) - 1;
        uint32_t set = sh.fixed ? slot : slot * sh.W + w;
        uint32_t key = set * sh.B + mag;
        uint32_t pos = atomicAdd(&cursor[key], 1u);
        uint32_t r = sh.fixed ? (w * sh.table_stride + ref) : ref;
        entries[offs[key] + pos] = r | (d < 0 ? 0x80000000u : 0u);
    }
}
__global__ void k_task_fill(const uint32_t *__restrict__ toffs, uint32_t *__restrict__ task_key, uint32_t nkeys) {
    uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nkeys) return;
    uint32_t t0 = toffs[key], t1 = toffs[key + 1];
    for (uint32_t t = t0; t < t1; t++) task_key[t] = key;
}

// ---------------------------------------------------------------- task order: longest first, equal lengths together
// Bucket sizes are Poisson distributed, so the tasks of 32 neighbouring buckets differ in length by +-30 % and a warp
// would wait for its longest task. A counting sort of the task ids by length (<= 64, block-aggregated atomics) makes
// every warp run tasks of (almost) one length; the partial of task t still goes to slot t, so nothing downstream changes.
#define BBP_MAX_S 64
__device__ __forceinline__ uint32_t task_len(const uint32_t *offs, const uint32_t *toffs, const uint32_t *task_key, uint32_t t, uint32_t S) {
    uint32_t key = task_key[t];
    uint32_t start = offs[key] + (t - toffs[key]) * S;
    return min(start + S, offs[key + 1]) - start;
}
// bins[0..65): number of tasks per length
__global__ void __launch_bounds__(256) k_task_hist(const uint32_t *__restrict__ offs, const uint32_t *__restrict__ toffs, const uint32_t *__restrict__ task_key,
                                                   uint32_t *__restrict__ bins, uint32_t nkeys, uint32_t S) {
    __shared__ uint32_t s_cnt[BBP_MAX_S + 1];
    if (threadIdx.x <= BBP_MAX_S) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < toffs[nkeys]) atomicAdd(&s_cnt[task_len(offs, toffs, task_key, t, S)], 1u);
    __syncthreads();
    if (threadIdx.x <= BBP_MAX_S && s_cnt[threadIdx.x]) atomicAdd(&bins[threadIdx.x], s_cnt[threadIdx.x]);
}
// bins[128 + len] = first position of length len in the permutation (descending length)
__global__ void k_task_bin_scan(uint32_t *__restrict__ bins) {
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int len = BBP_MAX_S; len >= 0; len--) { bins[128 + len] = acc; acc += bins[len]; }
    }
}
__global__ void __launch_bounds__(256) k_task_perm(const uint32_t *__restrict__ offs, const uint32_t *__restrict__ toffs, const uint32_t *__restrict__ task_key,
                                                   uint32_t *__restrict__ bins, uint32_t *__restrict__ perm, uint32_t nkeys, uint32_t S) {
    __shared__ uint32_t s_cnt[BBP_MAX_S + 1], s_base[BBP_MAX_S + 1];
    if (threadIdx.x <= BBP_MAX_S) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = t < toffs[nkeys];
    uint32_t len = 0, rank = 0;
    if (live) { len = task_len(offs, toffs, task_key, t, S); rank = atomicAdd(&s_cnt[len], 1u); }
    __syncthreads();
    if (threadIdx.x <= BBP_MAX_S && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&bins[128 + threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (live) perm[s_base[len] + rank] = t;
}

// ---------------------------------------------------------------- bucket accumulation: one thread per task
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_accumulate(const uint8_t *__restrict__ table, const uint32_t *__restrict__ entries,
                                                    const uint32_t *__restrict__ offs, const uint32_t *__restrict__ toffs,
                                                    const uint32_t *__restrict__ task_key, const uint32_t *__restrict__ perm,
                                                    uint8_t *__restrict__ partial, msm_shape sh) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n_tasks = toffs[sh.nkeys];
    if (i >= n_tasks) return;
    uint32_t t = perm[i];
    uint32_t key = task_key[t];
    uint32_t start = offs[key] + (t - toffs[key]) * sh.S;
    uint32_t end = min(start + sh.S, offs[key + 1]);
    uint32_t e = entries[start];
    ge acc = ge_from_niels(niels_load_ro(table + 96 * (size_t)(e & 0x7fffffffu)), (e >> 31) != 0);
    if (start + 1 < end) {
        // software pipeline: the gather of entry j+1 is in flight while entry j is being added
        e = entries[start + 1];
        niels q = niels_load_ro(table + 96 * (size_t)(e & 0x7fffffffu));
        for (uint32_t j = start + 2; j < end; j++) {
            uint32_t en = entries[j];
            niels qn = niels_load_ro(table + 96 * (size_t)(en & 0x7fffffffu));
            acc = ge_madd(acc, q, (e >> 31) != 0);
            q = qn; e = en;
        }
        acc = ge_madd(acc, q, (e >> 31) != 0);
    }
    ge_store(partial + 128 * (size_t)t, acc);
}

// ---------------------------------------------------------------- bucket reduction: sum_b b * bucket_b per bucket set
// Multi-level, every level one thread per group of <= 8 elements. An element stands for a block of `len` consecutive
// buckets and carries R = sum of the block's buckets and A = sum (local index, 1-based) * bucket. Merging g elements:
//   R' = sum_i R_i,   A' = sum_i A_i + len * sum_i i * R_i   (i = 0 .. g-1),   sum_i i R_i by running suffix sums.
// Level 1 reads the task partials of 8 buckets (len = 1, A_i = R_i = bucket_i); the last level leaves one element per
// bucket set whose A is the set total. Depth: log_8(buckets) levels of ~3*8 additions instead of one long chain.
#define BBP_RED_G 8
#define BBP_FOLD_MIN 8
// Buckets whose entries were split over MANY tasks (skewed digits: equal scalars, or a top window that only ever holds
// the recoding carry) are folded first by a whole warp: lanes sum a strided share of the task partials, then a 5-step
// tree through shared memory; the bucket sum replaces the first task's partial. Buckets with <= BBP_FOLD_MIN tasks (the
// common case) are left to level 1, which adds their few partials itself.
__global__ void __launch_bounds__(256) k_heavy_list(const uint32_t *__restrict__ toffs, uint32_t nkeys, uint32_t *__restrict__ list, uint32_t *__restrict__ count) {
    uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nkeys) return;
    if (toffs[key + 1] - toffs[key] > BBP_FOLD_MIN) list[atomicAdd(count, 1u)] = key;
}
#define BBP_FOLD_BLOCKS 1024
__global__ void __launch_bounds__(128) k_bucket_fold(uint8_t *__restrict__ partial, const uint32_t *__restrict__ toffs, const uint32_t *__restrict__ list,
                                                     const uint32_t *__restrict__ count) {
    __shared__ uint4 sm_u4[4 * 32 * 8];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *sm = (uint8_t *)sm_u4 + (size_t)warp * 32 * 128;
    const uint32_t n_heavy = *count;
    for (uint32_t h = blockIdx.x * 4 + warp; h < n_heavy; h += gridDim.x * 4) {
        uint32_t key = list[h];
        uint32_t t0 = toffs[key], t1 = toffs[key + 1];
        ge acc = ge_identity();
        bool nz = false;
        for (uint32_t t = t0 + lane; t < t1; t += 32) {
            ge p = ge_load(partial + 128 * (size_t)t);
            if (nz) acc = ge_add(acc, p);
            else { acc = p; nz = true; }
        }
        ge_store(sm + 128 * lane, acc);
        __syncwarp();
        for (uint32_t stride = 16; stride >= 1; stride >>= 1) {
            if (lane < stride) ge_store(sm + 128 * lane, ge_add(ge_load(sm + 128 * lane), ge_load(sm + 128 * (lane + stride))));
            __syncwarp();
        }
        if (lane == 0) ge_store(partial + 128 * (size_t)t0, ge_load(sm));
        __syncwarp();
    }
}
// level 1: groups of g buckets (task partials at toffs[bucket] ..; heavy buckets already folded into the first) -> (A, R)
__global__ void __launch_bounds__(128) k_reduce_level1(const uint8_t *__restrict__ partial, const uint32_t *__restrict__ toffs,
                                                       uint8_t *__restrict__ outA, uint8_t *__restrict__ outR, uint32_t n_groups, uint32_t g) {
    uint32_t ck = blockIdx.x * blockDim.x + threadIdx.x;
    if (ck >= n_groups) return;
    uint32_t k0 = ck * g;
    uint32_t to[BBP_RED_G + 1];
#pragma unroll
    for (uint32_t i = 0; i <= BBP_RED_G; i++) to[i] = (i <= g) ? toffs[k0 + i] : 0u;
    ge run = ge_identity(), acc = ge_identity();
    bool run_nz = false, acc_nz = false;
#pragma unroll 1
    for (uint32_t i = g; i-- > 0;) {
        uint32_t cnt = to[i + 1] - to[i];
        if (cnt > BBP_FOLD_MIN) cnt = 1;   // folded by k_bucket_fold into its first partial
        for (uint32_t t = to[i]; t < to[i] + cnt; t++) {
            ge p = ge_load(partial + 128 * (size_t)t);
            if (run_nz) run = ge_add(run, p);
            else { run = p; run_nz = true; }
        }
        if (run_nz) {
            if (acc_nz) acc = ge_add(acc, run);
            else { acc = run; acc_nz = true; }
        }
    }
    ge_store(outA + 128 * (size_t)ck, acc);
    ge_store(outR + 128 * (size_t)ck, run);
}
// merges g consecutive elements of block length `len` (a power of two)
__global__ void __launch_bounds__(128) k_reduce_merge(const uint8_t *__restrict__ inA, const uint8_t *__restrict__ inR, uint8_t *__restrict__ outA,
                                                      uint8_t *__restrict__ outR, uint32_t n_groups, uint32_t g, uint32_t len) {
    uint32_t ck = blockIdx.x * blockDim.x + threadIdx.x;
    if (ck >= n_groups) return;
    size_t e0 = (size_t)ck * g;
    // suffix running sums: run = R_{g-1} + ... + R_i ; w = sum_{i>=1} run_i = sum_i i * R_i
    ge run = ge_load(inR + 128 * (e0 + g - 1));
    ge w = run;
    ge A = ge_load(inA + 128 * (e0 + g - 1));
#pragma unroll 1
    for (uint32_t i = g - 1; i-- > 0;) {
        run = ge_add(run, ge_load(inR + 128 * (e0 + i)));
        if (i >= 1) w = ge_add(w, run);
        A = ge_add(A, ge_load(inA + 128 * (e0 + i)));
    }
    if (g > 1) {
        for (uint32_t m = len; m > 1; m >>= 1) w = ge_dbl(w);
        A = ge_add(A, w);
    }
    ge_store(outA + 128 * (size_t)ck, A);
    ge_store(outR + 128 * (size_t)ck, run);
}

// warp-cooperative merge of 32 consecutive elements (lane i holds element i): a suffix scan of R by shuffles (5 steps),
// then the two reductions sum_{i>=1} S_i = sum_i i R_i and sum_i A_i side by side (5 steps) — 10 dependent additions
// instead of the 3 * 32 of a serial walk. Used while a bucket set still has >= 32 elements.
BBP_DEV ge ge_shfl_down(const ge &p, uint32_t delta) {
    ge r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.X.v[i] = __shfl_down_sync(0xffffffffu, p.X.v[i], delta);
        r.Y.v[i] = __shfl_down_sync(0xffffffffu, p.Y.v[i], delta);
        r.Z.v[i] = __shfl_down_sync(0xffffffffu, p.Z.v[i], delta);
        r.T.v[i] = __shfl_down_sync(0xffffffffu, p.T.v[i], delta);
    }
    return r;
}
__global__ void __launch_bounds__(128) k_reduce_merge32(const uint8_t *__restrict__ inA, const uint8_t *__restrict__ inR, uint8_t *__restrict__ outA,
                                                        uint8_t *__restrict__ outR, uint32_t n_groups, uint32_t len) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t grp = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (grp >= n_groups) return;            // whole warps leave together
    size_t e = (size_t)grp * 32 + lane;
    ge S = ge_load(inR + 128 * e), A = ge_load(inA + 128 * e);
#pragma unroll 1
    for (uint32_t d = 1; d < 32; d <<= 1) {
        ge t = ge_shfl_down(S, d);
        if (lane + d < 32) S = ge_add(S, t);
    }
    // S = R_lane + ... + R_31 ; W accumulates the S of lanes >= 1
    ge W = (lane >= 1) ? S : ge_identity();
#pragma unroll 1
    for (uint32_t d = 16; d >= 1; d >>= 1) {
        ge tw = ge_shfl_down(W, d), ta = ge_shfl_down(A, d);
        if (lane < d) { W = ge_add(W, tw); A = ge_add(A, ta); }
    }
    if (lane == 0) {
        for (uint32_t m = len; m > 1; m >>= 1) W = ge_dbl(W);
        ge_store(outA + 128 * (size_t)grp, ge_add(A, W));
        ge_store(outR + 128 * (size_t)grp, S);
    }
}

// ---------------------------------------------------------------- final combine: four lanes per slot
// variable bases: Horner over the W window totals (c doublings per step); fixed bases: the single set total. The
// doublings are the only long dependency chain of the whole MSM ((W-1)*c of them), so each one is spread over four
// lanes: lane l owns coordinate l of the accumulator, the four squarings and the four products of dbl-2008-hwcd run
// side by side, operands travel by warp shuffles. Writes compressed (32 B) and / or extended (128 B) results.
struct ge4 {   // coordinate `lane & 3` of an extended point
    fe c;
};
BBP_DEV fe fe_shfl4(const fe &v, int src) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], src, 4);
    return r;
}
BBP_DEV fe ge4_finish(const fe &E, const fe &F, const fe &G, const fe &H, int l) {
    // lane 0: E*F (X), lane 1: G*H (Y), lane 2: F*G (Z), lane 3: E*H (T)
    fe m1 = (l == 0 || l == 3) ? E : (l == 1 ? G : F);
    fe m2 = (l == 0) ? F : ((l == 1 || l == 3) ? H : G);
    return fe_mul(m1, m2);
}
BBP_DEV fe ge4_dbl(const fe &c, int l) {
    fe X = fe_shfl4(c, 0), Y = fe_shfl4(c, 1);
    fe op = (l == 3) ? fe_add(X, Y) : c;
    fe sq = fe_sq(op);
    fe A = fe_shfl4(sq, 0), B = fe_shfl4(sq, 1), Zq = fe_shfl4(sq, 2), Sq = fe_shfl4(sq, 3);
    fe C = fe_dbl(Zq), D = fe_neg(A);
    fe E = fe_sub(fe_sub(Sq, A), B), G = fe_add(D, B), F = fe_sub(G, C), H = fe_sub(D, B);
    return ge4_finish(E, F, G, H, l);
}
// acc + q, q given in full on every lane
BBP_DEV fe ge4_add(const fe &c, const ge &q, int l) {
    fe X = fe_shfl4(c, 0), Y = fe_shfl4(c, 1), Z = fe_shfl4(c, 2), T = fe_shfl4(c, 3);
    fe a, b;
    if (l == 0) { a = fe_sub(Y, X); b = fe_sub(q.Y, q.X); }
    else if (l == 1) { a = fe_add(Y, X); b = fe_add(q.Y, q.X); }
    else if (l == 2) { a = T; b = fe_mul(q.T, fe_d2()); }
    else { a = fe_dbl(Z); b = q.Z; }
    fe pr = fe_mul(a, b);
    fe A = fe_shfl4(pr, 0), B = fe_shfl4(pr, 1), C = fe_shfl4(pr, 2), D = fe_shfl4(pr, 3);
    fe E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
    return ge4_finish(E, F, G, H, l);
}
__global__ void __launch_bounds__(32) k_combine(const uint8_t *__restrict__ set_total, uint8_t *__restrict__ out_ext,
                                                uint32_t *__restrict__ out_compressed, msm_shape sh) {
    const int l = threadIdx.x & 3;
    uint32_t slot = blockIdx.x * 8 + (threadIdx.x >> 2);
    bool live = slot < sh.n_slots;
    uint32_t sl = live ? slot : 0;   // idle lane groups shadow slot 0 so that the shuffles stay convergent
    fe c;
    if (sh.fixed) {
        c = fe_load(set_total + 128 * (size_t)sl + 32 * l);
    } else {
        const uint8_t *st = set_total + 128 * (size_t)sl * sh.W;
        c = fe_load(st + 128 * (size_t)(sh.W - 1) + 32 * l);
#pragma unroll 1
        for (uint32_t w = sh.W - 1; w-- > 0;) {
#pragma unroll 1
            for (uint32_t j = 0; j < sh.c; j++) c = ge4_dbl(c, l);
            c = ge4_add(c, ge_load(st + 128 * (size_t)w), l);
        }
    }
    ge acc;
    acc.X = fe_shfl4(c, 0); acc.Y = fe_shfl4(c, 1); acc.Z = fe_shfl4(c, 2); acc.T = fe_shfl4(c, 3);
    if (!live || l != 0) return;
    if (out_ext) ge_store(out_ext + 128 * (size_t)slot, acc);
    if (out_compressed) ge_compress_words(out_compressed + 8 * (size_t)slot, acc);
}

// ---------------------------------------------------------------- host-side engine
struct msm_engine {
    cudaStream_t stream = nullptr;
    // capacities
    size_t cap_pairs = 0, cap_keys = 0, cap_tasks = 0, cap_chunks = 0, cap_sets = 0;
    int32_t *digits = nullptr;
    uint32_t *hist = nullptr, *offs = nullptr, *cursor = nullptr, *toffs = nullptr, *entries = nullptr, *task_key = nullptr;
    uint32_t *tile_sums = nullptr, *total = nullptr, *task_perm = nullptr, *len_bins = nullptr;
    uint8_t *partial = nullptr, *lvl[4] = {nullptr, nullptr, nullptr, nullptr};   // lvl: ping-pong (A, R) arrays of the bucket reduction
    uint64_t launches = 0;
    // one-shot calls build the base table on a copy stream while recode / sort run here: the accumulation waits for this event
    cudaEvent_t table_ready = nullptr;
    // optional per-stage timing (bbp_set_profiling): events around recode / scans / scatter+fill / accumulate /
    // chunk reduce / window reduce / combine on the launching stream
    static const int N_STAGES = 7;
    bool profile = false;
    cudaEvent_t ev[N_STAGES + 1] = {};
    float stage_ms[N_STAGES] = {};
    bool stage_pending = false;

    int mark(int i) {
        if (!profile) return 0;
        if (!ev[i]) BBP_CUDA_OK(cudaEventCreate(&ev[i]));
        BBP_CUDA_OK(cudaEventRecord(ev[i], stream));
        return 0;
    }
    // blocks until the last profiled run has finished; ms[0..N_STAGES)
    int collect(float *ms) {
        if (stage_pending) {
            BBP_CUDA_OK(cudaEventSynchronize(ev[N_STAGES]));
            for (int i = 0; i < N_STAGES; i++) BBP_CUDA_OK(cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]));
            stage_pending = false;
        }
        for (int i = 0; i < N_STAGES; i++) ms[i] = stage_ms[i];
        return 0;
    }

    static size_t max_tasks(const msm_shape &sh) { return ((size_t)sh.n * sh.W + sh.S - 1) / sh.S + sh.nkeys; }

    void release() {
        cudaFree(digits); cudaFree(hist); cudaFree(offs); cudaFree(cursor); cudaFree(toffs); cudaFree(entries); cudaFree(task_key);
        cudaFree(tile_sums); cudaFree(total); cudaFree(partial); cudaFree(task_perm); cudaFree(len_bins);
        task_perm = len_bins = nullptr;
        for (int i = 0; i < 4; i++) { cudaFree(lvl[i]); lvl[i] = nullptr; }
        digits = nullptr; hist = offs = cursor = toffs = entries = task_key = tile_sums = total = nullptr;
        partial = nullptr;
        cap_pairs = cap_keys = cap_tasks = cap_chunks = cap_sets = 0;
        for (int i = 0; i <= N_STAGES; i++) if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; }
    }

    int reserve(const msm_shape &sh) {
        size_t pairs = (size_t)sh.n * sh.W, keys = sh.nkeys, tasks = max_tasks(sh), chunks = keys / 2 + 1, sets = (size_t)sh.n_slots * sh.sets_per_slot;
        if (pairs > cap_pairs) {
            cudaFree(digits); cudaFree(entries);
            BBP_CUDA_OK(cudaMalloc(&digits, pairs * 4));
            BBP_CUDA_OK(cudaMalloc(&entries, pairs * 4));
            cap_pairs = pairs;
        }
        if (keys > cap_keys) {
            cudaFree(hist); cudaFree(offs); cudaFree(cursor); cudaFree(toffs); cudaFree(tile_sums);
            BBP_CUDA_OK(cudaMalloc(&hist, (keys + 1) * 4));
            BBP_CUDA_OK(cudaMalloc(&offs, (keys + 1) * 4));
            BBP_CUDA_OK(cudaMalloc(&cursor, (keys + 1) * 4));
            BBP_CUDA_OK(cudaMalloc(&toffs, (keys + 1) * 4));
            BBP_CUDA_OK(cudaMalloc(&tile_sums, (keys / 4096 + 2) * 4));
            if (!total) BBP_CUDA_OK(cudaMalloc(&total, 4));
            cap_keys = keys;
        }
        if (tasks > cap_tasks) {
            cudaFree(task_key); cudaFree(partial); cudaFree(task_perm);
            BBP_CUDA_OK(cudaMalloc(&task_key, tasks * 4));
            BBP_CUDA_OK(cudaMalloc(&task_perm, tasks * 4));
            if (!len_bins) BBP_CUDA_OK(cudaMalloc(&len_bins, 256 * 4));
            BBP_CUDA_OK(cudaMalloc(&partial, tasks * 128));
            cap_tasks = tasks;
        }
        if (chunks > cap_chunks) {
            for (int i = 0; i < 4; i++) {
                cudaFree(lvl[i]);
                lvl[i] = nullptr;
                BBP_CUDA_OK(cudaMalloc(&lvl[i], chunks * 128));
            }
            cap_chunks = chunks;
        }
        cap_sets = sets;
        return 0;
    }

    int scan(const uint32_t *in, uint32_t *out, uint32_t n, uint32_t S) {
        uint32_t tiles = (n + 4095) / 4096;
        k_scan_tiles<<<tiles, 1024, 0, stream>>>(in, out, tile_sums, n, S);
        k_scan_tile_sums<<<1, 1024, 0, stream>>>(tile_sums, tiles, total);
        k_scan_add<<<(n + 1 + 255) / 256, 256, 0, stream>>>(out, tile_sums, n, total);
        launches += 3;
        return 0;
    }

    // window size heuristic (pairs per bucket set kept around 16-64)
    static msm_shape make_shape(uint32_t n, uint32_t n_per_slot, uint32_t base_mod, bool fixed, uint32_t table_c, uint32_t table_W, uint32_t table_stride) {
        msm_shape sh;
        sh.n = n; sh.n_per_slot = n_per_slot; sh.n_slots = (n + n_per_slot - 1) / n_per_slot; sh.base_mod = base_mod;
        sh.fixed = fixed ? 1 : 0;
        sh.ref_mode = 0; sh.colmap = nullptr; sh.colmap_len = 0; sh.grp_div = 1; sh.grp_stride = 0; sh.tail_ref = 0;
        if (fixed) {
            sh.c = table_c; sh.W = table_W; sh.table_stride = table_stride; sh.sets_per_slot = 1;
        } else {
            uint32_t c = 4;
            while (c < 16 && ((size_t)n_per_slot >> (c + 4)) >= 1) c++;   // >= 16 points per bucket on average
            if (const char *e = getenv("BBP_MSM_C")) { int v = atoi(e); if (v >= 4 && v <= 16) c = (uint32_t)v; }   // tuning knob
            sh.c = c; sh.W = 253 / c + 1; sh.table_stride = 0; sh.sets_per_slot = sh.W;
        }
        sh.B = 1u << (sh.c - 1);
        sh.nkeys = sh.n_slots * sh.sets_per_slot * sh.B;
        size_t pairs_per_set = fixed ? (size_t)n_per_slot * sh.W : n_per_slot;
        size_t avg = pairs_per_set / sh.B + 1;
        sh.S = (uint32_t)std::min<size_t>(std::max<size_t>(2 * avg, 16), 64);
        return sh;
    }

    // d_scalars: n x 32 B on device; d_table: niels table; outputs on device (either may be null).
    // phase 0 = whole pipeline; 1 = scalar side only (recode, sort, task table: needs no base table); 2 = the rest. A one-shot
    // call enqueues phase 1, then uploads and converts its points on the copy stream, then enqueues phase 2.
    int run(const msm_shape &sh, const uint8_t *d_scalars, const uint8_t *d_table, uint8_t *d_out_ext, uint8_t *d_out_compressed, int phase = 0) {
        if (sh.n == 0) return -1;
        size_t mt = max_tasks(sh);
        if (phase != 2) {
        int rc = reserve(sh);
        if (rc) return rc;
        BBP_CUDA_OK(cudaMemsetAsync(hist, 0, ((size_t)sh.nkeys + 1) * 4, stream));
        BBP_CUDA_OK(cudaMemsetAsync(cursor, 0, ((size_t)sh.nkeys + 1) * 4, stream));
        BBP_CUDA_OK(cudaMemsetAsync(len_bins, 0, 256 * 4, stream));
        if (mark(0)) return -100;
        k_recode<<<(sh.n + 127) / 128, 128, 0, stream>>>((const uint32_t *)d_scalars, digits, hist, sh);
        if (mark(1)) return -100;
        scan(hist, offs, sh.nkeys, 0);
        scan(hist, toffs, sh.nkeys, sh.S);
        if (mark(2)) return -100;
        k_scatter<<<(sh.n + 127) / 128, 128, 0, stream>>>(digits, offs, cursor, entries, sh);
        k_task_fill<<<(sh.nkeys + 255) / 256, 256, 0, stream>>>(toffs, task_key, sh.nkeys);
        k_task_hist<<<(unsigned)((mt + 255) / 256), 256, 0, stream>>>(offs, toffs, task_key, len_bins, sh.nkeys, sh.S);
        k_task_bin_scan<<<1, 32, 0, stream>>>(len_bins);
        k_task_perm<<<(unsigned)((mt + 255) / 256), 256, 0, stream>>>(offs, toffs, task_key, len_bins, task_perm, sh.nkeys, sh.S);
        launches += 3;
        if (mark(3)) return -100;
        }
        if (phase == 1) return 0;
        if (table_ready) BBP_CUDA_OK(cudaStreamWaitEvent(stream, table_ready, 0));
        {
            // resident CTAs per SM for the accumulation kernel (register budget 65536 / (128 * MINB)); BBP_ACC_MINB overrides
            static const int minb = [] { const char *e = getenv("BBP_ACC_MINB"); return e ? atoi(e) : 4; }();
            unsigned grid = (unsigned)((mt + 127) / 128);
            if (minb <= 3) k_accumulate<3><<<grid, 128, 0, stream>>>(d_table, entries, offs, toffs, task_key, task_perm, partial, sh);
            else if (minb == 4) k_accumulate<4><<<grid, 128, 0, stream>>>(d_table, entries, offs, toffs, task_key, task_perm, partial, sh);
            else if (minb == 5) k_accumulate<5><<<grid, 128, 0, stream>>>(d_table, entries, offs, toffs, task_key, task_perm, partial, sh);
            else k_accumulate<6><<<grid, 128, 0, stream>>>(d_table, entries, offs, toffs, task_key, task_perm, partial, sh);
        }
        if (mark(4)) return -100;
        // bucket reduction: B buckets per set -> 1 element per set, groups of <= 8 per level
        uint32_t per_set = sh.B, g = std::min<uint32_t>(BBP_RED_G, per_set);
        uint32_t sets = sh.n_slots * sh.sets_per_slot;
        uint32_t n_groups = sets * (per_set / g);
        // cursor[] is free after the scatter: [0, nkeys) becomes the list of heavy buckets, cursor[nkeys] (zeroed above) its length
        k_heavy_list<<<(sh.nkeys + 255) / 256, 256, 0, stream>>>(toffs, sh.nkeys, cursor, cursor + sh.nkeys);
        k_bucket_fold<<<BBP_FOLD_BLOCKS, 128, 0, stream>>>(partial, toffs, cursor, cursor + sh.nkeys);
        launches++;
        k_reduce_level1<<<(n_groups + 127) / 128, 128, 0, stream>>>(partial, toffs, lvl[0], lvl[1], n_groups, g);
        launches += 2;
        if (mark(5)) return -100;
        per_set /= g;
        uint32_t len = g;
        int cur = 0;
        while (per_set > 1) {
            if (per_set >= 32) {   // a warp per 32 elements
                g = 32;
                n_groups = sets * (per_set / g);
                k_reduce_merge32<<<(n_groups + 3) / 4, 128, 0, stream>>>(lvl[cur], lvl[cur + 1], lvl[cur ^ 2], lvl[(cur ^ 2) + 1], n_groups, len);
            } else {               // the last few elements: one thread walks them
                g = per_set;
                n_groups = sets;
                k_reduce_merge<<<(n_groups + 127) / 128, 128, 0, stream>>>(lvl[cur], lvl[cur + 1], lvl[cur ^ 2], lvl[(cur ^ 2) + 1], n_groups, g, len);
            }
            launches++;
            per_set /= g;
            len *= g;
            cur ^= 2;
        }
        if (mark(6)) return -100;
        k_combine<<<(sh.n_slots + 7) / 8, 32, 0, stream>>>(lvl[cur], d_out_ext, (uint32_t *)d_out_compressed, sh);
        if (mark(7)) return -100;
        if (profile) stage_pending = true;
        launches += 5;
        BBP_CUDA_OK(cudaGetLastError());
        return 0;
    }
};

}  // namespace bbp
