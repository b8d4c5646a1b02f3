// ORACLE — test infrastructure only (see fe.h header).
// Multiscalar multiplication sum_i s_i * P_i. The compressed result is algorithm-independent
// (SURVEY.md §8 a-10), so the oracle offers (1) a naive double-and-add sum used to pin everything
// else and (2) a Pippenger bucket method following what dalek's VartimeMultiscalarMul does for
// >= 190 points (signed radix-2^w digits, 2^(w-1) buckets, running-sum bucket
// reduction, columns high -> low) which is the CPU baseline that gets timed. Optional std::thread
// sharding over point ranges stands in for "all host cores" (dalek itself is single-threaded).
#pragma once
#include "ge.h"
#include <thread>
#include <vector>

namespace orc {

static inline ge msm_naive(const sc *s, const ge *p, size_t n) {
    ge acc = ge_identity();
    for (size_t i = 0; i < n; i++) acc = ge_add(acc, ge_scalarmul(s[i], p[i]));
    return acc;
}

// signed radix-2^w recoding of a canonical scalar (< 2^253): digits in [-2^(w-1), 2^(w-1)]
static inline void sc_to_radix_2w(std::vector<int32_t> &digits, const sc &s, int w) {
    int nd = (256 + w - 1) / w + 1;
    digits.assign(nd, 0);
    int64_t carry = 0;
    const int64_t radix = 1LL << w, half = radix >> 1;
    for (int i = 0; i < nd; i++) {
        int bit = i * w;
        uint64_t chunk = 0;
        if (bit < 256) {
            int word = bit >> 6, off = bit & 63;
            chunk = s.v[word] >> off;
            if (off + w > 64 && word < 3) chunk |= s.v[word + 1] << (64 - off);
            chunk &= (uint64_t)(radix - 1);
        }
        int64_t coef = (int64_t)chunk + carry;
        carry = (coef + half) >> w;
        digits[i] = (int32_t)(coef - (carry << w));
    }
}

static inline int pippenger_window(size_t n) {
    if (n < 500) return 6;
    if (n < 800) return 7;
    if (n < 4096) return 8;
    if (n < 32768) return 10;
    if (n < 262144) return 12;
    return 14;
}

static inline ge msm_pippenger_serial(const sc *s, const ge *p, size_t n) {
    if (n == 0) return ge_identity();
    const int w = pippenger_window(n);
    const size_t nb = (size_t)1 << (w - 1);
    std::vector<std::vector<int32_t>> dig(n);
    size_t nd = 0;
    for (size_t i = 0; i < n; i++) { sc_to_radix_2w(dig[i], s[i], w); nd = dig[i].size(); }
    std::vector<ge> buckets(nb);
    ge total = ge_identity();
    for (size_t col = nd; col-- > 0;) {
        for (int k = 0; k < w; k++) total = ge_dbl(total);
        for (size_t b = 0; b < nb; b++) buckets[b] = ge_identity();
        for (size_t i = 0; i < n; i++) {
            int32_t d = dig[i][col];
            if (d > 0) buckets[d - 1] = ge_add(buckets[d - 1], p[i]);
            else if (d < 0) buckets[-d - 1] = ge_sub(buckets[-d - 1], p[i]);
        }
        ge run = ge_identity(), sum = ge_identity();
        for (size_t b = nb; b-- > 0;) {
            run = ge_add(run, buckets[b]);
            sum = ge_add(sum, run);
        }
        total = ge_add(total, sum);
    }
    return total;
}

static inline ge msm_pippenger(const sc *s, const ge *p, size_t n, int threads) {
    if (threads <= 1 || n < 1024) return msm_pippenger_serial(s, p, n);
    std::vector<ge> part(threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        size_t lo = n * t / threads, hi = n * (t + 1) / threads;
        th.emplace_back([&, t, lo, hi] { part[t] = msm_pippenger_serial(s + lo, p + lo, hi - lo); });
    }
    for (auto &x : th) x.join();
    ge acc = ge_identity();
    for (int t = 0; t < threads; t++) acc = ge_add(acc, part[t]);
    return acc;
}

}  // namespace orc
