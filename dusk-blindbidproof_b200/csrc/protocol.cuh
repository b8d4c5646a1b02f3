// Batched blind-bid prover and verifier on one B200 (host orchestration; kernels in sc_kernels.cuh / msm.cuh /
// fixedbase.cuh). Mirrors Proof::prove (src/blindbid/proof.rs:36-91) and Verify::verify (src/blindbid/verify.rs:47-89)
// including everything bulletproofs does underneath them (SURVEY.md §8 a-4 .. a-8), for a whole batch of requests at
// once: the reference serves one request per pool thread (src/futures/main.rs:52-62), here one launch serves them all.
//
// Division of labour
//   host (one thread per proof, std::thread)  Merlin transcripts, TranscriptRng draws, witness evaluation, (de)serialisation
//   GPU                                       every group operation and every O(n) scalar vector operation
// There is no CPU implementation of the group: without the kernels nothing here can produce a point.
//
// RNG contract: callers pass the commitment blindings and the 32 "external" bytes that the reference
// takes from thread_rng inside TranscriptRngBuilder::finalize; with those fixed every proof byte is deterministic.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <chrono>
#include <cstdlib>
#include <map>
#include <memory>
#include <thread>
#include "circuit.h"
#include "ctx.cuh"
#include "fixedbase.cuh"
#include "rng_kernels.cuh"
#include "sc_kernels.cuh"
#include "small_msm.cuh"

namespace bbp {

// ---------------------------------------------------------------- small host utilities
// host threads for the per-proof phases: BBP_HOST_THREADS, else the hardware threads divided among the processes that
// share this host (torchrun exports LOCAL_WORLD_SIZE: one process per GPU)
inline size_t host_threads() {
    static const size_t hw = [] {
        if (const char *e = getenv("BBP_HOST_THREADS")) return (size_t)std::max(1, atoi(e));
        size_t h = std::max<size_t>(std::thread::hardware_concurrency(), 1);
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) h = std::max<size_t>(h / (size_t)std::max(1, atoi(e)), 1);
        return h;
    }();
    return hw;
}
// BBP_KECCAK_THREAD=1 selects the one-thread-per-sponge Keccak kernels (k_rng_draws, k_verify_transcript) instead of the
// warp-cooperative ones; read per call so that the tests can compare both
inline bool keccak_per_thread() { const char *e = getenv("BBP_KECCAK_THREAD"); return e && atoi(e) != 0; }
// Persistent host worker pool (one per process): the per-proof phases are short (tens of microseconds per proof), so
// spawning threads per phase would cost as much as the phase. Workers sleep on a condition variable between phases.
class host_pool {
  public:
    static host_pool &get() {
        static host_pool p;
        return p;
    }
    // runs fn(lo, hi) over a partition of [0, n) into contiguous chunks, at most one chunk per worker; blocks until done.
    // Calls from different threads (several contexts) are serialised.
    template <class F>
    void run_chunks(size_t n, F fn) {
        if (n == 0) return;
        size_t nt = std::min<size_t>(workers_.size() + 1, n);
        if (nt <= 1) { fn((size_t)0, n); return; }
        std::lock_guard<std::mutex> call_lock(call_mu_);
        std::function<void(size_t)> job = [&](size_t t) {
            size_t lo = n * t / nt, hi = n * (t + 1) / nt;
            if (lo < hi) fn(lo, hi);
        };
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &job; n_chunks_ = nt; next_ = 1; pending_ = nt - 1; generation_++;
        }
        cv_.notify_all();
        job(0);   // the caller takes chunk 0
        // help with leftover chunks, then wait for the workers
        for (;;) {
            size_t t;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (next_ >= n_chunks_) break;
                t = next_++;
            }
            job(t);
            std::lock_guard<std::mutex> lk(mu_);
            pending_--;
        }
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
    }

  private:
    host_pool() {
        size_t n = host_threads();
        for (size_t i = 1; i < n; i++) workers_.emplace_back([this] { loop(); });
    }
    ~host_pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || (generation_ != seen && job_ && next_ < n_chunks_); });
            if (stop_) return;
            seen = generation_;
            while (job_ && next_ < n_chunks_) {
                size_t t = next_++;
                std::function<void(size_t)> *job = job_;
                lk.unlock();
                (*job)(t);
                lk.lock();
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    std::function<void(size_t)> *job_ = nullptr;
    size_t n_chunks_ = 0, next_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

template <class F>
inline void parallel_chunks(size_t n, F fn) { host_pool::get().run_chunks(n, fn); }
template <class F>
inline void parallel_for(size_t n, F fn) {
    parallel_chunks(n, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) fn(i); });
}

// Montgomery's trick: xs[i] <- xs[i]^-1 for all i with ONE exponentiation (inputs must be non-zero, which a Fiat-Shamir
// challenge is except with probability 2^-252; a zero would poison the chunk, so it falls back to single inversions)
inline void sc_batch_invert(sc *xs, size_t n) {
    if (n == 0) return;
    std::vector<sc> pre(n);
    sc acc = sc_one();
    bool has_zero = false;
    for (size_t i = 0; i < n; i++) { has_zero = has_zero || sc_iszero(xs[i]); pre[i] = acc; acc = sc_mul(acc, xs[i]); }
    if (has_zero) { for (size_t i = 0; i < n; i++) xs[i] = sc_invert(xs[i]); return; }
    sc inv = sc_invert(acc);
    for (size_t i = n; i-- > 0;) { sc t = sc_mul(inv, pre[i]); inv = sc_mul(inv, xs[i]); xs[i] = t; }
}

// BBP_TRACE=1: per-phase wall-clock of the batched prover / verifier on stderr (each phase ends with a stream sync)
struct phase_trace {
    bool on;
    const char *what;
    std::chrono::steady_clock::time_point t0;
    std::vector<std::pair<const char *, double>> marks;
    explicit phase_trace(const char *w) : on(getenv("BBP_TRACE") != nullptr), what(w), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *name) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        marks.push_back({name, std::chrono::duration<double, std::milli>(t1 - t0).count()});
        t0 = t1;
    }
    ~phase_trace() {
        if (!on) return;
        double tot = 0;
        for (auto &m : marks) tot += m.second;
        fprintf(stderr, "[bbp trace] %s total %.3f ms:", what, tot);
        for (auto &m : marks) fprintf(stderr, " %s=%.3f", m.first, m.second);
        fprintf(stderr, "\n");
    }
};

// BBP_TRACE=1: device timeline of one stream-ordered sequence (events between kernels, read after the final sync)
struct event_timeline {
    bool on;
    std::vector<std::pair<const char *, cudaEvent_t>> ev;
    event_timeline() : on(getenv("BBP_TRACE") != nullptr) {}
    void mark(const char *name, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, st);
        ev.push_back({name, e});
    }
    void report(const char *what) {   // call after the streams have been synchronised
        if (!on || ev.empty()) return;
        fprintf(stderr, "[bbp timeline] %s (ms since %s):", what, ev[0].first);
        for (size_t i = 1; i < ev.size(); i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[0].second, ev[i].second);
            fprintf(stderr, " %s=%.3f", ev[i].first, ms);
        }
        fprintf(stderr, "\n");
        for (auto &e : ev) cudaEventDestroy(e.second);
        ev.clear();
    }
};

struct dev_buf {
    uint8_t *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4;
        if (cudaMalloc(&p, want) != cudaSuccess) { fprintf(stderr, "bbp: cudaMalloc(%zu) failed\n", want); return BBP_ERR_CUDA; }
        cap = want;
        return 0;
    }
    void release() { cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// grow-only pinned host staging (fresh pageable vectors of tens of MB cost more in page faults than the copy itself)
struct host_buf {
    uint8_t *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4;
        if (cudaHostAlloc((void **)&p, want, cudaHostAllocDefault) != cudaSuccess) { fprintf(stderr, "bbp: cudaHostAlloc(%zu) failed\n", want); return BBP_ERR_CUDA; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct dev_template {
    std::shared_ptr<const circuit_template> tpl;
    uint32_t *row_ptr = nullptr, *entries = nullptr, *const_j = nullptr, *const_idx = nullptr;
    uint32_t *coef = nullptr;      // generic circuits only: coefficient index per entry (sc_kernels.cuh: flatten_term)
    uint32_t n_long = 0, long_rows[BBP_MAX_LONG] = {};   // the longest CSR rows among wL / wR / wO (sc_kernels.cuh: flatten_long_rows)
};

// window table over the generators: 24 windows of 11 bits (264 >= 254 bits). BBP_WT_C (10 .. 14) is a tuning knob, read once.
static const uint32_t WT_C = [] {
    const char *e = getenv("BBP_WT_C");
    int v = e ? atoi(e) : 11;
    return (uint32_t)((v >= 10 && v <= 14) ? v : 11);
}();
static const uint32_t WT_W = (253 + WT_C) / WT_C;
static const uint32_t WT2_C = 6, WT2_W = 43;   // second table, small windows: 32 buckets per slot, for the many tiny slots of the
                                               // IPP materialisation MSM (16 scalars each)
static const uint32_t IPP_NF = 128;            // hybrid IPP: folded bases are materialised when the vectors reach this length

struct proto_state {
    std::map<uint64_t, dev_template> templates;
    uint8_t *comb = nullptr;       // Pedersen comb table
    uint8_t *wtable = nullptr;     // generator window table, WT_W rows of n_gens niels entries
    uint8_t *wtable2 = nullptr;    // same with WT2_C-bit windows
    uint32_t *colmap = nullptr;    // compact slot -> generator column map of the materialisation MSM
    uint32_t colmap_n = 0, colmap_gcols = 0;
    uint8_t *dtable = nullptr;     // digit-multiple table of the latency path (small_msm.cuh), built on first use
    dev_buf sm_partial, lane_partials, shard_gather, pub_shared;
    uint32_t *ipp_colmap = nullptr;   // early IPP rounds, compact slots: lg n maps of 2 (1 + n) generator columns (L slot, R slot)
    uint32_t ipp_colmap_n = 0, ipp_colmap_gcols = 0;
    dev_buf fext, ftab;            // materialised folded bases: extended, then niels (+ B at the tail)
    dev_buf chal, zpow, ypow, yinvpow, wit, vbl, blind3, poly, tout, a, b, sG, sH, slots, ab, pub, dyn_sc, dyn_pts, dyn_niels, stat, stat_red,
        msm_out, msm_ext, flags, valid, commit_in, commit_out, rng_states, rng_raw, bw_digests, rp_blobs, rp_chal0, rp_dyn0;
    host_buf h_wit, h_states;
    uint32_t resident_B = 0, resident_ds = 0;   // requests / dynamic points per request of the last combined verification pass (verify_regroup_pass)
    cudaStream_t rng_stream = nullptr;   // the device TranscriptRng chain runs beside the A_I1 / A_O1 commitments
    cudaEvent_t ev_up = nullptr, ev_rng = nullptr, ev_dyn = nullptr, ev_head = nullptr, ev_pow = nullptr;
    int proof_versioned = 1;       // R1CSProof::to_bytes layout (SURVEY.md §8c risk R1): 1 = leading phase byte, 0 = legacy 14-point form
};

inline int proto_side_stream(proto_state *ps) {
    if (!ps->rng_stream) {
        // highest priority: what runs here are latency chains of few warps (the RNG chain, the variable-base MSM of a
        // verification); their blocks should take the first free slots beside the wide scalar kernels of the main stream
        int prio_lo = 0, prio_hi = 0;
        BBP_CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        BBP_CUDA_OK(cudaStreamCreateWithPriority(&ps->rng_stream, cudaStreamNonBlocking, prio_hi));
        BBP_CUDA_OK(cudaEventCreateWithFlags(&ps->ev_up, cudaEventDisableTiming));
        BBP_CUDA_OK(cudaEventCreateWithFlags(&ps->ev_rng, cudaEventDisableTiming));
        BBP_CUDA_OK(cudaEventCreateWithFlags(&ps->ev_dyn, cudaEventDisableTiming));
        BBP_CUDA_OK(cudaEventCreateWithFlags(&ps->ev_head, cudaEventDisableTiming));
        BBP_CUDA_OK(cudaEventCreateWithFlags(&ps->ev_pow, cudaEventDisableTiming));
    }
    return 0;
}

inline proto_state *proto_get(bbp_ctx *ctx) {
    if (!ctx->proto) ctx->proto = new proto_state();
    return ctx->proto;
}

void proto_release(proto_state *ps) {
    if (!ps) return;
    for (auto &kv : ps->templates) { cudaFree(kv.second.row_ptr); cudaFree(kv.second.entries); cudaFree(kv.second.const_j); cudaFree(kv.second.const_idx); cudaFree(kv.second.coef); }
    cudaFree(ps->comb); cudaFree(ps->wtable); cudaFree(ps->wtable2); cudaFree(ps->colmap); cudaFree(ps->ipp_colmap); cudaFree(ps->dtable);
    ps->sm_partial.release(); ps->lane_partials.release(); ps->shard_gather.release(); ps->pub_shared.release();
    ps->fext.release(); ps->ftab.release();
    dev_buf *all[] = {&ps->chal, &ps->zpow, &ps->ypow, &ps->yinvpow, &ps->wit, &ps->vbl, &ps->blind3, &ps->poly, &ps->tout, &ps->a, &ps->b, &ps->sG, &ps->sH,
                      &ps->slots, &ps->ab, &ps->pub, &ps->dyn_sc, &ps->dyn_pts, &ps->dyn_niels, &ps->stat, &ps->stat_red, &ps->msm_out, &ps->msm_ext,
                      &ps->flags, &ps->valid, &ps->commit_in, &ps->commit_out, &ps->rng_states, &ps->rng_raw, &ps->bw_digests, &ps->rp_blobs, &ps->rp_chal0, &ps->rp_dyn0};
    for (dev_buf *b : all) b->release();
    ps->h_wit.release(); ps->h_states.release();
    if (ps->rng_stream) cudaStreamDestroy(ps->rng_stream);
    if (ps->ev_up) cudaEventDestroy(ps->ev_up);
    if (ps->ev_rng) cudaEventDestroy(ps->ev_rng);
    if (ps->ev_dyn) cudaEventDestroy(ps->ev_dyn);
    if (ps->ev_head) cudaEventDestroy(ps->ev_head);
    if (ps->ev_pow) cudaEventDestroy(ps->ev_pow);
    delete ps;
}

// one-time tables that need the generator set of the context
inline int proto_tables(bbp_ctx *ctx) {
    proto_state *ps = proto_get(ctx);
    if (ctx->n_gens < 2) return BBP_ERR_INVALID_GENERATORS_LENGTH;
    if (!ps->pub_shared.p) {   // public values every blind-bid proof shares: 1 and the 90 MiMC round constants (circuit.h: PV_ONE, PV_MIMC)
        std::vector<sc> sh(PV_SEED);
        sh[PV_ONE] = sc_one();
        for (uint32_t i = 0; i < MIMC_ROUNDS; i++) sh[PV_MIMC + i] = mimc_constants()[i];
        int prc = ps->pub_shared.ensure(sh.size() * 32);
        if (prc) return prc;
        BBP_CUDA_OK(cudaMemcpyAsync(ps->pub_shared.p, sh.data(), sh.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
        BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));   // the source is a local
    }
    if (!ps->comb) {
        BBP_CUDA_OK(cudaMalloc(&ps->comb, (size_t)BBP_COMB_ENTRIES * 96));
        k_build_comb<<<1, 128, 0, ctx->stream>>>(ctx->d_gens_ext, ps->comb);
        ctx->launches++;
    }
    if (!ps->wtable && ctx->n_gens > 2) {
        BBP_CUDA_OK(cudaMalloc(&ps->wtable, (size_t)WT_W * ctx->n_gens * 96));
        k_build_window_table<<<(unsigned)((ctx->n_gens + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_gens_ext, ps->wtable, (uint32_t)ctx->n_gens, WT_C, WT_W,
                                                                                              (uint32_t)ctx->n_gens);
        ctx->launches++;
    }
    if (!ps->wtable2 && ctx->n_gens > 2) {
        BBP_CUDA_OK(cudaMalloc(&ps->wtable2, (size_t)WT2_W * ctx->n_gens * 96));
        k_build_window_table<<<(unsigned)((ctx->n_gens + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_gens_ext, ps->wtable2, (uint32_t)ctx->n_gens, WT2_C, WT2_W,
                                                                                              (uint32_t)ctx->n_gens);
        ctx->launches++;
    }
    BBP_CUDA_OK(cudaGetLastError());
    return 0;
}

inline int template_upload(bbp_ctx *ctx, dev_template &dt);
inline void template_free(dev_template &dt);
inline int proto_template(bbp_ctx *ctx, uint32_t n_commit, uint32_t n_toggle, dev_template **out) {
    proto_state *ps = proto_get(ctx);
    uint64_t key = ((uint64_t)n_commit << 32) | n_toggle;
    auto it = ps->templates.find(key);
    if (it == ps->templates.end()) {
        if (ps->templates.size() >= TEMPLATE_CACHE_MAX) {   // keyed by request-chosen counts: bounded (no call holds a pointer across calls)
            BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
            for (auto &kv : ps->templates) template_free(kv.second);
            ps->templates.clear();
        }
        dev_template dt;
        dt.tpl = blindbid_template(n_commit, n_toggle);
        int urc = template_upload(ctx, dt);
        if (urc) return urc;
        it = ps->templates.emplace(key, dt).first;
    }
    *out = &it->second;
    return 0;
}
inline void template_free(dev_template &dt) {
    cudaFree(dt.row_ptr); cudaFree(dt.entries); cudaFree(dt.const_j); cudaFree(dt.const_idx); cudaFree(dt.coef);
    dt.row_ptr = dt.entries = dt.const_j = dt.const_idx = dt.coef = nullptr;
}
// uploads dt.tpl's CSR; fills the device pointers and the long-row list
inline int template_upload(bbp_ctx *ctx, dev_template &dt) {
    {
        {
        const circuit_template &t = *dt.tpl;
        BBP_CUDA_OK(cudaMalloc(&dt.row_ptr, t.row_ptr.size() * 4));
        BBP_CUDA_OK(cudaMalloc(&dt.entries, std::max<size_t>(t.entries.size(), 1) * 4));
        BBP_CUDA_OK(cudaMalloc(&dt.const_j, std::max<size_t>(t.const_j.size(), 1) * 4));
        BBP_CUDA_OK(cudaMalloc(&dt.const_idx, std::max<size_t>(t.const_idx.size(), 1) * 4));
        BBP_CUDA_OK(cudaMemcpyAsync(dt.row_ptr, t.row_ptr.data(), t.row_ptr.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        BBP_CUDA_OK(cudaMemcpyAsync(dt.entries, t.entries.data(), t.entries.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        BBP_CUDA_OK(cudaMemcpyAsync(dt.const_j, t.const_j.data(), t.const_j.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        BBP_CUDA_OK(cudaMemcpyAsync(dt.const_idx, t.const_idx.data(), t.const_idx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (!t.coef.empty()) {
            BBP_CUDA_OK(cudaMalloc(&dt.coef, t.coef.size() * 4));
            BBP_CUDA_OK(cudaMemcpyAsync(dt.coef, t.coef.data(), t.coef.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        {
            std::vector<std::pair<uint32_t, uint32_t>> lens;   // (length, row) of rows above the threshold, longest first
            for (uint32_t r = 0; r < 3 * t.n1; r++) {
                uint32_t len = t.row_ptr[r + 1] - t.row_ptr[r];
                if (len > BBP_LONG_ROW) lens.emplace_back(len, r);
            }
            std::sort(lens.rbegin(), lens.rend());
            dt.n_long = (uint32_t)std::min<size_t>(lens.size(), BBP_MAX_LONG);
            for (uint32_t k = 0; k < dt.n_long; k++) dt.long_rows[k] = lens[k].second;
        }
        }
    }
    return 0;
}

// k_powers / k_verify_scalars take their small tables in dynamic shared memory; circuits beyond the blind-bid sizes can need
// more than the 48 KB a kernel gets without opting in
inline int sc_kernels_smem_opt_in(uint32_t q, uint32_t n, uint32_t lg) {
    static std::mutex mu;
    static size_t cur_p = 0, cur_v = 0;
    std::lock_guard<std::mutex> lock(mu);
    const size_t p = k_powers_smem(q, n), v = k_verify_scalars_smem(n, lg);
    if (p > cur_p && p + 8192 > 48 * 1024) {
        BBP_CUDA_OK(cudaFuncSetAttribute(k_powers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p));
        cur_p = p;
    }
    if (v > cur_v && v + 20480 > 48 * 1024) {
        BBP_CUDA_OK(cudaFuncSetAttribute(k_verify_scalars, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v));
        cur_v = v;
    }
    return 0;
}
inline uint32_t next_pow2_u32(uint32_t n) { uint32_t p = 1; while (p < n) p <<= 1; return p; }
inline uint32_t log2_u32(uint32_t n) { uint32_t l = 0; while ((1u << l) < n) l++; return l; }

inline int h2d(bbp_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!bytes) return 0;
    BBP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
// Waiting for the context's stream. cudaStreamSynchronize spins on a core; with several processes sharing the host (one per
// GPU, each with a few lane threads waiting most of the time) the spinning threads take the cores the host phases of the
// other ranks may need; BBP_BLOCKING_SYNC=1 makes the wait sleep on a blocking-sync event instead (a wake-up of ~20 us per
// wait). Off by default: measured on 8 GPUs / 32 cores it changes nothing (3.19 M proofs/s either way at 1024 per call).
inline bool blocking_sync() {
    static const bool on = [] {
        const char *e = getenv("BBP_BLOCKING_SYNC");
        return e && atoi(e) != 0;
    }();
    return on;
}
inline int wait_stream(bbp_ctx *ctx) {
    if (blocking_sync() && ctx->ev_block) {
        BBP_CUDA_OK(cudaEventRecord(ctx->ev_block, ctx->stream));
        BBP_CUDA_OK(cudaEventSynchronize(ctx->ev_block));
    } else {
        BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}
inline int d2h_sync(bbp_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes) BBP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return wait_stream(ctx);
}

// n commitments v*B + r*B_blinding; vals = n x (value, blinding) reduced scalars on the host; out = n x 32 B on the host
inline bool small_msm_ok(size_t total_terms);
inline bool ct_commit();
inline int msm_gens_small(bbp_ctx *ctx, const sc *d_scalars, uint32_t slot_len, uint32_t n_slots, const uint32_t *colmap, uint32_t colmap_slots,
                          uint8_t *d_out_compressed, uint8_t *d_out_ext, bool uniform);
inline int pedersen_commit_host(bbp_ctx *ctx, const sc *vals, size_t n, uint8_t *out) {
    proto_state *ps = proto_get(ctx);
    int rc;
    if (n == 0) return 0;   // a circuit without commitments
    if ((rc = ps->commit_in.ensure(n * 64)) || (rc = ps->commit_out.ensure(n * 32))) return rc;
    if ((rc = h2d(ctx, ps->commit_in.p, vals, n * 64))) return rc;
    if (ct_commit() || (n <= 64 && small_msm_ok(2 * n))) {
        // a handful of commitments (one request's V or T points): the comb kernel is a 128-addition chain per thread
        // (~0.4 ms); as two-column slots [B, B_blinding] of the latency path the additions spread over four warps.
        // BBP_CT_COMMIT: every batch size, uniform instruction stream (values and blindings are secret)
        if ((rc = msm_gens_small(ctx, ps->commit_in.as<sc>(), 2, (uint32_t)n, nullptr, 0, ps->commit_out.p, nullptr, ct_commit()))) return rc;
        return d2h_sync(ctx, out, ps->commit_out.p, n * 32);
    }
    k_pedersen_commit<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ps->commit_in.as<sc>(), ps->comb, ps->commit_out.as<uint32_t>(), (uint32_t)n);
    ctx->launches++;
    return d2h_sync(ctx, out, ps->commit_out.p, n * 32);
}

// n_slots MSMs over the generator window table; scalars: n_slots x slot_len on the device (slot_len <= n_gens, column i
// of a slot multiplies generator i). Results: compressed (n_slots x 32 B) and / or extended (n_slots x 128 B), on device.
// Small problems take the latency path (small_msm.cuh): two launches instead of the engine's ~20. The crossover is in
// total terms (BBP_SMALL_MSM_MAX, default 66000 = the IPP rounds of up to 16 proofs, the commitments of up to 5; measured:
// 6.0 / 6.9 / 7.9 / 9.1 ms per prove call at batch 1 / 2 / 4 / 8 against 9.3 / 9.6 / 10.3 / 11.6 through the engine, equal
// from 16 on; 0 disables the path).
inline bool small_msm_ok(size_t total_terms) {
    const char *e = getenv("BBP_SMALL_MSM_MAX");
    return total_terms <= (size_t)(e ? atol(e) : 66000);
}
// BBP_CT_COMMIT=1: MSMs whose scalars are SECRET (the V / T Pedersen commitments, A_I1 / A_O1 / S1 — what the reference
// computes with dalek's constant-time multiscalar_mul) take the digit-table path with a uniform
// instruction stream whatever their size, instead of the variable-time bucket engine. Read per call.
inline bool ct_commit() { const char *e = getenv("BBP_CT_COMMIT"); return e && atoi(e) != 0; }
inline int msm_gens_small(bbp_ctx *ctx, const sc *d_scalars, uint32_t slot_len, uint32_t n_slots, const uint32_t *colmap, uint32_t colmap_slots,
                          uint8_t *d_out_compressed, uint8_t *d_out_ext, bool uniform = false) {
    proto_state *ps = proto_get(ctx);
    if (!ps->dtable) {
        const size_t entries = (size_t)SM_W * ctx->n_gens;
        BBP_CUDA_OK(cudaMalloc(&ps->dtable, entries * SM_D * 96));
        k_build_digit_table<<<(unsigned)((entries + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_gens_ext, ps->dtable, (uint32_t)ctx->n_gens);
        ctx->launches++;
    }
    const uint32_t chunks = (slot_len + SM_THREADS / SM_GROUPS - 1) / (SM_THREADS / SM_GROUPS);
    int rc;
    if ((rc = ps->sm_partial.ensure((size_t)n_slots * chunks * 128))) return rc;
    k_small_msm_partial<<<dim3(chunks, n_slots), SM_THREADS, 0, ctx->stream>>>(d_scalars, slot_len, ps->dtable, (uint32_t)ctx->n_gens, colmap,
                                                                                 colmap_slots ? colmap_slots : 1, ps->sm_partial.p, uniform ? 1u : 0u);
    k_small_msm_final<<<n_slots, SM_THREADS, 0, ctx->stream>>>(ps->sm_partial.p, chunks, d_out_ext, d_out_compressed);
    ctx->launches += 2;
    BBP_CUDA_OK(cudaGetLastError());
    return 0;
}

inline int msm_gens_device(bbp_ctx *ctx, const sc *d_scalars, uint32_t slot_len, uint32_t n_slots, uint8_t *d_out_compressed, uint8_t *d_out_ext,
                           bool secret = false) {
    proto_state *ps = proto_get(ctx);
    if (slot_len > ctx->n_gens || !ps->wtable) return BBP_ERR_INVALID_GENERATORS_LENGTH;
    if (secret && ct_commit()) return msm_gens_small(ctx, d_scalars, slot_len, n_slots, nullptr, 0, d_out_compressed, d_out_ext, true);
    if (small_msm_ok((size_t)n_slots * slot_len)) return msm_gens_small(ctx, d_scalars, slot_len, n_slots, nullptr, 0, d_out_compressed, d_out_ext);
    msm_shape sh = msm_engine::make_shape(n_slots * slot_len, slot_len, slot_len, true, WT_C, WT_W, (uint32_t)ctx->n_gens);
    return ctx->msm.run(sh, (const uint8_t *)d_scalars, ps->wtable, d_out_ext, d_out_compressed);
}

// Inner-product argument for a batch of P proofs whose a, b, sG, sH live in SB (set up by k_ipp_init / k_rp_ipp_init).
// Early rounds: L_j, R_j as MSMs over the ORIGINAL generators with product-form scalars (no folding, see sc_kernels.cuh).
// Once the vectors have shrunk to IPP_NF the folded bases are materialised ONCE (one MSM with 2*IPP_NF tiny slots per
// proof over the small-window table), and the remaining rounds are 2*IPP_NF+1-point MSMs over those per-proof bases.
// on_round(j, lr) receives the compressed L_j, R_j (64 B per proof) and must fill CH_UJ / CH_UJINV of every proof in chal.
template <class F>
inline int ipp_rounds(bbp_ctx *ctx, sc_batch &SB, uint32_t P, std::vector<sc> &chal, F on_round, phase_trace &trace) {
    proto_state *ps = proto_get(ctx);
    const uint32_t n = SB.n, lg = SB.lg_n, gcols = SB.gcols, slot_len = 2 + 2 * gcols;
    const char *hyb_env = getenv("BBP_IPP_HYBRID");   // 0 = never, 2 = always (tests), default: batches of >= 32
    const int hyb = hyb_env ? atoi(hyb_env) : 1;
    const uint32_t shard_G = ctx->shard_world > 1 ? ctx->shard_world : (ctx->shard_emulate > 1 ? (uint32_t)ctx->shard_emulate : 1);
    const bool sharded = shard_G > 1;
    if (sharded && ctx->shard_world > 1 && !ctx->shard_allgather) return BBP_ERR_INPUT;
    // sharded: every round's L_j / R_j stay MSMs over the ORIGINAL generators, so that the column partition never moves
    const bool hybrid = !sharded && hyb && n >= 4 * IPP_NF && (P >= 32 || hyb == 2);   // small batches are launch bound: fewer, larger rounds win
    const char *nf_env = getenv("BBP_IPP_NF");   // tuning knob (power of two, 16 .. n / 4)
    const uint32_t nf = nf_env ? (uint32_t)atoi(nf_env) : IPP_NF;
    const uint32_t j0 = hybrid ? lg - log2_u32(nf) : lg;
    int rc;
    std::vector<uint8_t> lr((size_t)P * 64);
    SB.fac_n = n; SB.late = 0;
    const char *cp_env = getenv("BBP_IPP_COMPACT");
    SB.compact = !sharded && (cp_env ? atoi(cp_env) : 1) && n >= 4;
    SB.shard_G = shard_G; SB.shard_g = ctx->shard_world > 1 ? ctx->shard_rank : 0;
    if (sharded && ((rc = ps->msm_ext.ensure((size_t)2 * P * 128)) || (rc = ps->shard_gather.ensure((size_t)shard_G * 2 * P * 128)))) return rc;
    const uint32_t cslot = 1 + n;
    if (SB.compact && (ps->ipp_colmap_n != n || ps->ipp_colmap_gcols != gcols)) {
        // round j, slot s (0 = L, 1 = R), entry e: e = 0 -> B; then n/2 G columns and n/2 H columns, block by block
        std::vector<uint32_t> cm((size_t)lg * 2 * cslot);
        for (uint32_t j = 0; j < lg; j++) {
            const uint32_t nj = n >> j, nh = nj >> 1, half = n >> 1;
            for (uint32_t s = 0; s < 2; s++) {
                uint32_t *o = &cm[((size_t)j * 2 + s) * cslot];
                o[0] = 0;
                for (uint32_t e = 0; e < n; e++) {
                    uint32_t fam = e >= half, tt = e % half, blk = tt / nh, off = tt % nh, lo = blk * nj + off, hi = lo + nh;
                    uint32_t idx = (s == 0) ? (fam == 0 ? hi : lo) : (fam == 0 ? lo : hi);
                    o[1 + e] = 2 + fam * gcols + idx;
                }
            }
        }
        cudaFree(ps->ipp_colmap);
        ps->ipp_colmap = nullptr;
        BBP_CUDA_OK(cudaMalloc(&ps->ipp_colmap, cm.size() * 4));
        BBP_CUDA_OK(cudaMemcpyAsync(ps->ipp_colmap, cm.data(), cm.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        { int wr = wait_stream(ctx); if (wr) return wr; }
        ps->ipp_colmap_n = n; ps->ipp_colmap_gcols = gcols;
    }
    if ((rc = ps->msm_out.ensure((size_t)P * 64))) return rc;
    for (uint32_t j = 0; j < lg; j++) {
        uint32_t mode = 0;
        if (hybrid && j == j0) {
            // ---- materialise F_G, F_H: compact scalars -> MSM over the generator table -> per-proof niels tables (+ B)
            if (ps->colmap_n != n || ps->colmap_gcols != gcols) {
                std::vector<uint32_t> cm((size_t)2 * n);
                const uint32_t E = n / nf;
                for (uint32_t idx = 0; idx < 2 * n; idx++) {
                    uint32_t fam = idx / n, r = idx % n, k = r / E, e = r % E;
                    cm[idx] = 2 + fam * gcols + k + nf * e;
                }
                cudaFree(ps->colmap);
                ps->colmap = nullptr;
                BBP_CUDA_OK(cudaMalloc(&ps->colmap, cm.size() * 4));
                BBP_CUDA_OK(cudaMemcpyAsync(ps->colmap, cm.data(), cm.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
                { int wr = wait_stream(ctx); if (wr) return wr; }
                ps->colmap_n = n; ps->colmap_gcols = gcols;
            }
            const size_t n_f_pts = (size_t)P * 2 * nf;
            if ((rc = ps->fext.ensure(n_f_pts * 128)) || (rc = ps->ftab.ensure((n_f_pts + 1) * 96))) return rc;
            SB.mat = SB.slots;   // the slot buffer is free between rounds and large enough (2 slots of 2 + 2 gcols >= 2 n)
            k_ipp_materialize<<<P, BBP_SC_THREADS, 0, ctx->stream>>>(SB, j0);
            ctx->launches++;
            msm_shape sh = msm_engine::make_shape(P * 2 * n, n / nf, n / nf, true, WT2_C, WT2_W, (uint32_t)ctx->n_gens);
            sh.ref_mode = 1; sh.colmap = ps->colmap; sh.colmap_len = 2 * n;
            if ((rc = ctx->msm.run(sh, (const uint8_t *)SB.mat, ps->wtable2, ps->fext.p, nullptr))) return rc;
            size_t thr = (n_f_pts + BBP_NIELS_BATCH - 1) / BBP_NIELS_BATCH;
            k_ext_to_niels<<<(unsigned)((thr + 127) / 128), 128, 0, ctx->stream>>>(ps->fext.p, ps->ftab.p, (uint32_t)n_f_pts);
            ctx->launches++;
            BBP_CUDA_OK(cudaMemcpyAsync(ps->ftab.p + n_f_pts * 96, ctx->d_gens_niels, 96, cudaMemcpyDeviceToDevice, ctx->stream));
            SB.fac_n = nf; SB.late = 1;
            mode = 2;   // the fold of this round was applied by k_ipp_materialize
            if (trace.on) { cudaStreamSynchronize(ctx->stream); trace.mark("gpu_ipp_materialize"); }
        }
        k_ipp_round<<<P, BBP_SC_THREADS, 0, ctx->stream>>>(SB, j, mode);
        ctx->launches++;
        if (sharded) {
            // this rank's columns -> 2 P partial sums (extended); all ranks' partials -> gather buffer; local sum + compression.
            // a, b, the factors and c_L / c_R are replicated (128 KB per proof: cheaper to recompute than to exchange), so the only
            // traffic per round is 2 x 128 B per proof per GPU and every rank derives the same challenge from the same bytes
            const size_t part_bytes = (size_t)2 * P * 128;
            if (ctx->shard_world > 1) {
                if ((rc = msm_gens_device(ctx, SB.slots, slot_len, 2 * P, nullptr, ps->msm_ext.p))) return rc;
                if (ctx->shard_allgather(ctx->shard_user, ps->msm_ext.p, ps->shard_gather.p, part_bytes)) return BBP_ERR_NCCL;
            } else {
                for (uint32_t g = 0; g < shard_G; g++) {   // emulation: the shards one after the other on this GPU
                    if (g) {
                        SB.shard_g = g;
                        k_ipp_round<<<P, BBP_SC_THREADS, 0, ctx->stream>>>(SB, j, mode | 2);   // slots only: the fold was applied above
                        ctx->launches++;
                    }
                    if ((rc = msm_gens_device(ctx, SB.slots, slot_len, 2 * P, nullptr, ps->shard_gather.p + g * part_bytes))) return rc;
                }
                SB.shard_g = 0;
            }
            k_sum_ranks_compress<<<(2 * P + 63) / 64, 64, 0, ctx->stream>>>(ps->shard_gather.p, 2 * P, shard_G, ps->msm_out.as<uint32_t>());
            ctx->launches++;
        } else if (!SB.late && SB.compact && small_msm_ok((size_t)2 * P * cslot)) {
            if ((rc = msm_gens_small(ctx, SB.slots, cslot, 2 * P, ps->ipp_colmap + (size_t)j * 2 * cslot, 2, ps->msm_out.p, nullptr))) return rc;
        } else if (!SB.late && SB.compact) {
            msm_shape sh = msm_engine::make_shape(2 * P * cslot, cslot, cslot, true, WT_C, WT_W, (uint32_t)ctx->n_gens);
            sh.ref_mode = 1; sh.colmap = ps->ipp_colmap + (size_t)j * 2 * cslot; sh.colmap_len = 2 * cslot;
            if ((rc = ctx->msm.run(sh, (const uint8_t *)SB.slots, ps->wtable, nullptr, ps->msm_out.p))) return rc;
        } else if (!SB.late) {
            if ((rc = msm_gens_device(ctx, SB.slots, slot_len, 2 * P, ps->msm_out.p, nullptr))) return rc;
        } else {
            const uint32_t sl = 2 * nf + 1;
            msm_shape sh = msm_engine::make_shape(P * 2 * sl, sl, sl, false, 0, 0, 0);
            sh.ref_mode = 2; sh.grp_div = 2 * sl; sh.grp_stride = 2 * nf; sh.tail_ref = P * 2 * nf;
            if ((rc = ctx->msm.run(sh, (const uint8_t *)SB.slots, ps->ftab.p, nullptr, ps->msm_out.p))) return rc;
        }
        if ((rc = d2h_sync(ctx, lr.data(), ps->msm_out.p, lr.size()))) return rc;
        trace.mark(SB.late ? "gpu_ipp_round_late" : "gpu_ipp_round");
        on_round(j, lr);
        if ((rc = h2d(ctx, ps->chal.p, chal.data(), chal.size() * 32))) return rc;
        trace.mark("host_ipp_round");
    }
    k_ipp_round<<<P, BBP_SC_THREADS, 0, ctx->stream>>>(SB, lg, 1);
    ctx->launches++;
    return 0;
}

// ================================================================ R1CSProof (de)serialisation
struct r1cs_proof_host {
    uint8_t A_I1[32], A_O1[32], S1[32], A_I2[32], A_O2[32], S2[32], T_1[32], T_3[32], T_4[32], T_5[32], T_6[32];
    sc t_x, t_x_blinding, e_blinding;
    std::vector<uint8_t> LR;   // L_0 R_0 L_1 R_1 ...
    sc a, b;
};

inline bool all_zero32(const uint8_t *p) { uint8_t acc = 0; for (int i = 0; i < 32; i++) acc |= p[i]; return acc == 0; }

// R1CSProof::to_bytes (call site src/blindbid/proof.rs:125)
inline std::vector<uint8_t> r1cs_to_bytes(const r1cs_proof_host &p, bool versioned) {
    std::vector<uint8_t> out;
    auto put = [&](const uint8_t *b) { out.insert(out.end(), b, b + 32); };
    auto puts = [&](const sc &s) { uint8_t t[32]; sc_tobytes(t, s); put(t); };
    bool one_phase = all_zero32(p.A_I2) && all_zero32(p.A_O2) && all_zero32(p.S2);
    if (versioned) out.push_back(one_phase ? 0 : 1);
    put(p.A_I1); put(p.A_O1); put(p.S1);
    if (!versioned || !one_phase) { put(p.A_I2); put(p.A_O2); put(p.S2); }
    put(p.T_1); put(p.T_3); put(p.T_4); put(p.T_5); put(p.T_6);
    puts(p.t_x); puts(p.t_x_blinding); puts(p.e_blinding);
    out.insert(out.end(), p.LR.begin(), p.LR.end());
    puts(p.a); puts(p.b);
    return out;
}

// R1CSProof::from_bytes (call site src/blindbid/proof.rs:154); false = FormatError
inline bool r1cs_from_bytes(r1cs_proof_host &p, const uint8_t *in, size_t len, bool versioned) {
    int version = 1;
    if (versioned) {
        if (len < 1) return false;
        version = in[0];
        in++; len--;
    }
    if (len % 32 != 0) return false;
    size_t minlen;
    if (version == 0) minlen = 11 * 32;
    else if (version == 1) minlen = 14 * 32;
    else return false;
    if (len < minlen) return false;
    size_t pos = 0;
    auto get = [&](uint8_t *b) { memcpy(b, in + pos, 32); pos += 32; };
    get(p.A_I1); get(p.A_O1); get(p.S1);
    if (version == 0) { memset(p.A_I2, 0, 32); memset(p.A_O2, 0, 32); memset(p.S2, 0, 32); }
    else { get(p.A_I2); get(p.A_O2); get(p.S2); }
    get(p.T_1); get(p.T_3); get(p.T_4); get(p.T_5); get(p.T_6);
    if (!sc_from_canonical(p.t_x, in + pos)) return false;
    pos += 32;
    if (!sc_from_canonical(p.t_x_blinding, in + pos)) return false;
    pos += 32;
    if (!sc_from_canonical(p.e_blinding, in + pos)) return false;
    pos += 32;
    // InnerProductProof::from_bytes
    size_t rest = len - pos, ne = rest / 32;
    if (ne < 2 || (ne - 2) % 2 != 0) return false;
    size_t lg_n = (ne - 2) / 2;
    if (lg_n >= 32) return false;
    p.LR.assign(in + pos, in + pos + 64 * lg_n);
    if (!sc_from_canonical(p.a, in + pos + 64 * lg_n)) return false;
    if (!sc_from_canonical(p.b, in + pos + 64 * lg_n + 32)) return false;
    return true;
}

// ================================================================ prover
struct prove_job {
    // inputs (all reduced / parsed by the caller of prove_batch)
    sc d, k, y, y_inv, q, z_img, seed;
    std::vector<sc> pub_list;       // L items (Scalar::from_bits semantics applied)
    uint64_t toggle = 0;
    std::vector<sc> blindings;      // 4 + L
    uint8_t rng_seed[32];
    // outputs
    int status = 0;
    std::vector<uint8_t> proof;
    std::vector<uint8_t> commitments, t_c;   // 4 x 32, L x 32
};

// What the R1CS prover core needs from its caller, per proof bi < B. All proofs of a call share one circuit template.
// Two callers: the blind-bid driver (Proof::prove, src/blindbid/proof.rs:36-91: the circuit of src/gadgets.rs, witness by
// the evaluator or by k_blindbid_witness) and the generic bulletproofs surface (bbp_r1cs_prove: circuit and witness
// recorded by the caller through ConstraintSystem::{multiply, constrain}).
struct prove_source {
    uint32_t B = 0;
    dev_template *dt = nullptr;
    // committed values and their blindings (m each), the 32 external RNG bytes (RNG contract)
    std::function<void(size_t, std::vector<sc> &, std::vector<sc> &, const uint8_t *&)> inputs;
    // a fresh transcript positioned where Prover::new leaves it (label, then "r1cs v1"); the core owns it
    std::function<merlin_transcript *(size_t)> transcript;
    // host witness: writes a_L, a_R, a_O (n1 scalars each) given the committed values
    std::function<void(size_t, const std::vector<sc> &, sc *, sc *, sc *)> witness;
    // optional device witness of the blind-bid circuit (bb_L = list length, 0 = not available): fills
    // (d, k, y_inv, seed, items[bb_L]) and the toggle index; false = this request cannot take the device path
    uint32_t bb_L = 0;
    std::function<bool(size_t, sc *, uint32_t *)> bb_witness_in;
    // result: V (m x 32 B compressed), the proof bytes, the transcript as the proof leaves it
    std::function<void(size_t, const uint8_t *, std::vector<uint8_t> &&, const merlin_transcript &)> done;
    std::function<void(int)> fail_all;
};

// Prover::prove for S.B proofs over one circuit (bulletproofs 1.0.4 @ 4a05305 r1cs/prover.rs; call site src/blindbid/proof.rs:88)
inline int prove_core(bbp_ctx *ctx, prove_source &S) {
    proto_state *ps = proto_get(ctx);
    const uint32_t B = S.B;
    int rc;
    if ((rc = proto_tables(ctx))) return rc;
    dev_template *dt = S.dt;
    const circuit_template &T = *dt->tpl;
    const uint32_t n1 = T.n1, m = T.m, n = next_pow2_u32(n1), lg = log2_u32(n);
    if (n > ctx->gens_capacity || ctx->party_capacity < 1) {
        S.fail_all(BBP_ERR_INVALID_GENERATORS_LENGTH);   // bp_gens.gens_capacity < padded_n
        return 0;
    }
    const uint32_t gcols = ctx->gens_capacity * ctx->party_capacity, slot_len = 2 + 2 * gcols;
    const uint32_t L = S.bb_L;

    struct hstate {
        std::unique_ptr<merlin_transcript> tr;
        std::unique_ptr<merlin_rng> rng;
        std::vector<sc> v, bl;
        const uint8_t *rng_seed = nullptr;
        r1cs_proof_host pf;
        sc i_bl, o_bl, s_bl, tbl[6];   // tbl[1] (t_2 blinding) comes from the device
    };
    std::vector<hstate> hs(B);
    phase_trace trace("prove_group");

    // ---- phase 0: witness + V commitments
    // the 2 n1 blinding-vector draws run on the device for large batches (one warp per proof continues the transcript RNG;
    // the chain is a few ms of latency whatever the batch; host threads draw ~2.4 ms per proof each, so the device wins
    // once the batch exceeds about two proofs per host thread)
    const char *rng_env = getenv("BBP_DEVICE_RNG_MIN_BATCH");   // read per call so that tests can force either path
    const int rng_threshold = rng_env ? atoi(rng_env) : (int)(2 * host_threads());
    const bool device_rng = (int)B >= rng_threshold;
    // witness staging (pinned, grow-only), vector-major [a_L | a_R | a_O | s_L | s_R], each B x n1: the evaluator writes in place
    const size_t wit_count = (size_t)B * (device_rng ? 3 : 5) * n1;
    if ((rc = ps->h_wit.ensure(wit_count * 32))) return rc;
    sc *wit = ps->h_wit.as<sc>();
    // large batches of the blind-bid circuit evaluate the witness on the device too (one thread per proof,
    // k_blindbid_witness); the host path (the caller's evaluator, or its recorded assignments) serves the rest
    bool device_witness = device_rng && L > 0 && (bool)S.bb_witness_in;
    std::vector<sc> commit_vals((size_t)B * m * 2), wit_in(device_witness ? (size_t)B * (4 + L) : 0);
    std::vector<uint32_t> toggles_u32(device_witness ? B : 0);
    std::atomic<bool> all_device{true};
    parallel_for(B, [&](size_t bi) {
        hstate &H = hs[bi];
        S.inputs(bi, H.v, H.bl, H.rng_seed);
        for (uint32_t i = 0; i < m; i++) { commit_vals[((size_t)bi * m + i) * 2] = H.v[i]; commit_vals[((size_t)bi * m + i) * 2 + 1] = H.bl[i]; }
        if (device_witness && !S.bb_witness_in(bi, &wit_in[bi * (4 + L)], &toggles_u32[bi])) all_device = false;
    });
    if (device_witness && !all_device) device_witness = false;
    if (!device_witness)
        parallel_for(B, [&](size_t bi) {
            S.witness(bi, hs[bi].v, &wit[((size_t)0 * B + bi) * n1], &wit[((size_t)1 * B + bi) * n1], &wit[((size_t)2 * B + bi) * n1]);
        });
    trace.mark("host_witness");
    std::vector<uint8_t> V((size_t)B * m * 32);
    if ((rc = pedersen_commit_host(ctx, commit_vals.data(), (size_t)B * m, V.data()))) return rc;
    trace.mark("gpu_V_commit");

    // ---- phase 1: transcript up to the blinding draws; upload witness
    if (device_rng) {
        if ((rc = ps->h_states.ensure((size_t)B * BBP_STROBE_STATE_BYTES))) return rc;
        if (!ps->rng_stream) {
            if ((rc = proto_side_stream(ps))) return rc;
        }
    }
    uint8_t *rng_states = ps->h_states.p;
    std::vector<sc> vbl((size_t)B * m), blind3((size_t)B * 3);
    parallel_for(B, [&](size_t bi) {
        hstate &H = hs[bi];
        H.tr.reset(S.transcript(bi));                                 // Transcript::new(label), then Prover::new's "r1cs v1"
        for (uint32_t i = 0; i < m; i++) H.tr->append_point("V", &V[((size_t)bi * m + i) * 32]);
        H.tr->append_u64("m", m);
        merlin_rng_builder rb = H.tr->build_rng();
        for (uint32_t i = 0; i < m; i++) {
            uint8_t b[32];
            sc_tobytes(b, H.bl[i]);
            rb.rekey_with_witness_bytes("v_blinding", b, 32);
            vbl[(size_t)bi * m + i] = H.bl[i];
        }
        H.rng.reset(new merlin_rng(rb.finalize(H.rng_seed)));
        H.i_bl = H.rng->random_scalar(); H.o_bl = H.rng->random_scalar(); H.s_bl = H.rng->random_scalar();
        blind3[bi * 3] = H.i_bl; blind3[bi * 3 + 1] = H.o_bl; blind3[bi * 3 + 2] = H.s_bl;
        auto W = [&](uint32_t k) { return &wit[((size_t)k * B + bi) * n1]; };
        if (device_rng) {
            H.rng->export_state(&rng_states[bi * BBP_STROBE_STATE_BYTES]);
        } else {
            for (uint32_t i = 0; i < n1; i++) W(3)[i] = H.rng->random_scalar();   // s_L
            for (uint32_t i = 0; i < n1; i++) W(4)[i] = H.rng->random_scalar();   // s_R
        }
    });

    trace.mark("host_rng_draws");
    // device buffers
    if ((rc = ps->chal.ensure((size_t)B * CH_N * 32)) || (rc = ps->zpow.ensure((size_t)B * T.q * 32)) || (rc = ps->ypow.ensure((size_t)B * n * 32)) ||
        (rc = ps->yinvpow.ensure((size_t)B * n * 32)) || (rc = ps->wit.ensure((size_t)B * 5 * n1 * 32)) || (rc = ps->vbl.ensure((size_t)B * m * 32)) ||
        (rc = ps->blind3.ensure((size_t)B * 96)) || (rc = ps->poly.ensure((size_t)B * 6 * n1 * 32)) || (rc = ps->tout.ensure((size_t)B * 8 * 32)) ||
        (rc = ps->a.ensure((size_t)B * n * 32)) || (rc = ps->b.ensure((size_t)B * n * 32)) || (rc = ps->sG.ensure((size_t)B * n * 32)) ||
        (rc = ps->sH.ensure((size_t)B * n * 32)) || (rc = ps->slots.ensure((size_t)B * 3 * slot_len * 32)) || (rc = ps->ab.ensure((size_t)B * 64)) ||
        (rc = ps->msm_out.ensure((size_t)B * 3 * 32)))
        return rc;
    if (device_witness) {
        if ((rc = ps->pub.ensure(wit_in.size() * 32 + 90 * 32 + (size_t)B * 4))) return rc;
        sc *d_in = ps->pub.as<sc>(), *d_c = d_in + wit_in.size();
        uint32_t *d_tg = (uint32_t *)(d_c + 90);
        if ((rc = h2d(ctx, d_in, wit_in.data(), wit_in.size() * 32)) || (rc = h2d(ctx, d_c, mimc_constants().data(), 90 * 32)) ||
            (rc = h2d(ctx, d_tg, toggles_u32.data(), (size_t)B * 4)))
            return rc;
        sc *dwit = ps->wit.as<sc>();
        k_blindbid_witness<<<(B + 31) / 32, 32, 0, ctx->stream>>>(d_in, d_tg, d_c, B, L, n1, dwit, dwit + (size_t)B * n1, dwit + (size_t)2 * B * n1);
        ctx->launches++;
    } else if ((rc = h2d(ctx, ps->wit.p, wit, wit_count * 32))) return rc;
    if ((rc = h2d(ctx, ps->vbl.p, vbl.data(), vbl.size() * 32)) ||
        (rc = h2d(ctx, ps->blind3.p, blind3.data(), blind3.size() * 32)))
        return rc;
    if (device_rng) {
        const size_t sb = (size_t)B * BBP_STROBE_STATE_BYTES;
        if ((rc = ps->rng_states.ensure(sb)) || (rc = h2d(ctx, ps->rng_states.p, rng_states, sb))) return rc;
        // the draw chain (one thread per proof, latency bound) runs on its own stream next to the A_I1 / A_O1 MSMs
        BBP_CUDA_OK(cudaEventRecord(ps->ev_up, ctx->stream));
        BBP_CUDA_OK(cudaStreamWaitEvent(ps->rng_stream, ps->ev_up, 0));
        if ((rc = ps->rng_raw.ensure((size_t)B * 2 * n1 * 64))) return rc;
        if (keccak_per_thread())
            k_rng_draws<<<(B + 31) / 32, 32, 0, ps->rng_stream>>>(ps->rng_states.p, B, 2 * n1, ps->rng_raw.as<uint32_t>());
        else   // one warp per proof
            k_rng_draws_warp<<<(B + 3) / 4, 128, 0, ps->rng_stream>>>(ps->rng_states.p, B, 2 * n1, ps->rng_raw.as<uint32_t>());
        k_wide_reduce<<<(unsigned)(((size_t)B * 2 * n1 + 127) / 128), 128, 0, ps->rng_stream>>>(ps->rng_raw.as<uint32_t>(), B, 2 * n1, n1, (size_t)B * n1,
                                                                                                 ps->wit.as<sc>() + (size_t)3 * B * n1);
        ctx->launches += 2;
        BBP_CUDA_OK(cudaMemcpyAsync(rng_states, ps->rng_states.p, sb, cudaMemcpyDeviceToHost, ps->rng_stream));
        BBP_CUDA_OK(cudaEventRecord(ps->ev_rng, ps->rng_stream));
    }
    sc_batch SB;
    memset(&SB, 0, sizeof SB);
    SB.n_proofs = B; SB.n1 = n1; SB.q = T.q; SB.m = m; SB.n = n; SB.lg_n = lg; SB.n_pub = T.n_pub; SB.gcols = gcols;
    SB.row_ptr = dt->row_ptr; SB.entries = dt->entries; SB.const_j = dt->const_j; SB.const_idx = dt->const_idx; SB.n_const = (uint32_t)T.const_j.size();
    SB.n_long = dt->n_long;
    memcpy(SB.long_rows, dt->long_rows, sizeof SB.long_rows);
    SB.chal = ps->chal.as<sc>(); SB.zpow = ps->zpow.as<sc>(); SB.ypow = ps->ypow.as<sc>(); SB.yinvpow = ps->yinvpow.as<sc>();
    SB.coef = dt->coef;
    if (dt->coef) {   // generic circuit: its coefficient table, once (shared by the batch: n_pub = 0 -> stride 0)
        if ((rc = ps->pub.ensure(T.coef_table.size() * 32)) || (rc = h2d(ctx, ps->pub.p, T.coef_table.data(), T.coef_table.size() * 32))) return rc;
        SB.pub = ps->pub.as<sc>();
        SB.n_pub = 0;
    }
    const sc *dw = ps->wit.as<sc>();
    const size_t vs = (size_t)B * n1;
    SB.aL = dw; SB.aR = dw + vs; SB.aO = dw + 2 * vs; SB.sL = dw + 3 * vs; SB.sR = dw + 4 * vs;
    SB.vbl = ps->vbl.as<sc>(); SB.blind3 = ps->blind3.as<sc>(); SB.poly = ps->poly.as<sc>(); SB.tout = ps->tout.as<sc>();
    SB.a = ps->a.as<sc>(); SB.b = ps->b.as<sc>(); SB.sG = ps->sG.as<sc>(); SB.sH = ps->sH.as<sc>(); SB.slots = ps->slots.as<sc>(); SB.ab_out = ps->ab.as<sc>();

    // ---- phase 2 (GPU): A_I1, A_O1, S1
    k_commit_slots<<<2 * B, BBP_SC_THREADS, 0, ctx->stream>>>(SB, 0, 2, 0);
    ctx->launches++;
    if ((rc = msm_gens_device(ctx, SB.slots, slot_len, 2 * B, ps->msm_out.p, nullptr, true))) return rc;   // witness scalars: secret
    if (device_rng) BBP_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ps->ev_rng, 0));
    k_commit_slots<<<B, BBP_SC_THREADS, 0, ctx->stream>>>(SB, 2, 1, 2 * B);
    ctx->launches++;
    if ((rc = msm_gens_device(ctx, SB.slots + (size_t)2 * B * slot_len, slot_len, B, ps->msm_out.p + (size_t)2 * B * 32, nullptr, true))) return rc;
    std::vector<uint8_t> pts((size_t)B * 3 * 32);
    if ((rc = d2h_sync(ctx, pts.data(), ps->msm_out.p, pts.size()))) return rc;
    trace.mark("gpu_A_S_msm");

    // ---- phase 3 (host): y, z
    std::vector<sc> chal((size_t)B * CH_N, sc_zero());
    parallel_chunks(B, [&](size_t lo_, size_t hi_) {
    for (size_t bi = lo_; bi < hi_; bi++) {
        hstate &H = hs[bi];
        r1cs_proof_host &P = H.pf;
        if (device_rng) H.rng->import_state(&rng_states[bi * BBP_STROBE_STATE_BYTES]);
        memcpy(P.A_I1, &pts[(bi * 2) * 32], 32); memcpy(P.A_O1, &pts[(bi * 2 + 1) * 32], 32); memcpy(P.S1, &pts[((size_t)2 * B + bi) * 32], 32);
        H.tr->append_point("A_I1", P.A_I1); H.tr->append_point("A_O1", P.A_O1); H.tr->append_point("S1", P.S1);
        H.tr->r1cs_1phase_domain_sep();     // the circuit has no randomised (second phase) constraints
        memset(P.A_I2, 0, 32); memset(P.A_O2, 0, 32); memset(P.S2, 0, 32);
        H.tr->append_point("A_I2", P.A_I2); H.tr->append_point("A_O2", P.A_O2); H.tr->append_point("S2", P.S2);
        sc *c = &chal[bi * CH_N];
        c[CH_Y] = H.tr->challenge_scalar("y");
        c[CH_Z] = H.tr->challenge_scalar("z");
        c[CH_YINV] = c[CH_Y];
    }
    // y^-1 for the whole chunk with one exponentiation (Montgomery's trick), like the u_j of the IPP rounds
    std::vector<sc> inv(hi_ - lo_);
    for (size_t bi = lo_; bi < hi_; bi++) inv[bi - lo_] = chal[bi * CH_N + CH_Y];
    sc_batch_invert(inv.data(), inv.size());
    for (size_t bi = lo_; bi < hi_; bi++) chal[bi * CH_N + CH_YINV] = inv[bi - lo_];
    });
    if ((rc = h2d(ctx, ps->chal.p, chal.data(), chal.size() * 32))) return rc;

    // ---- phase 4 (GPU): power tables, flattened constraints, l / r polynomials, t_1 .. t_6
    if ((rc = sc_kernels_smem_opt_in(SB.q, SB.n, SB.lg_n))) return rc;
    k_powers<<<B, BBP_SC_THREADS, k_powers_smem(SB.q, SB.n), ctx->stream>>>(SB);
    k_polys<<<B, BBP_SC_THREADS, 0, ctx->stream>>>(SB);
    ctx->launches += 2;
    std::vector<sc> tout((size_t)B * 8);
    if ((rc = d2h_sync(ctx, tout.data(), ps->tout.p, tout.size() * 32))) return rc;
    trace.mark("yz+gpu_polys");

    // ---- phase 5: T_1, T_3, T_4, T_5, T_6
    std::vector<sc> tvals((size_t)B * 5 * 2);
    parallel_for(B, [&](size_t bi) {
        hstate &H = hs[bi];
        static const int tk[5] = {0, 2, 3, 4, 5};   // t_1, t_3, t_4, t_5, t_6
        for (int k = 0; k < 5; k++) {
            H.tbl[tk[k]] = H.rng->random_scalar();
            tvals[(bi * 5 + k) * 2] = tout[bi * 8 + tk[k]];
            tvals[(bi * 5 + k) * 2 + 1] = H.tbl[tk[k]];
        }
        H.tbl[1] = tout[bi * 8 + 6];   // t_2 blinding = <wV, v_blinding>
    });
    std::vector<uint8_t> Tp((size_t)B * 5 * 32);
    if ((rc = pedersen_commit_host(ctx, tvals.data(), (size_t)B * 5, Tp.data()))) return rc;
    trace.mark("gpu_T_commit");

    // ---- phase 6 (host): u, x, t(x), blindings, w
    parallel_for(B, [&](size_t bi) {
        hstate &H = hs[bi];
        r1cs_proof_host &P = H.pf;
        memcpy(P.T_1, &Tp[(bi * 5) * 32], 32); memcpy(P.T_3, &Tp[(bi * 5 + 1) * 32], 32); memcpy(P.T_4, &Tp[(bi * 5 + 2) * 32], 32);
        memcpy(P.T_5, &Tp[(bi * 5 + 3) * 32], 32); memcpy(P.T_6, &Tp[(bi * 5 + 4) * 32], 32);
        H.tr->append_point("T_1", P.T_1); H.tr->append_point("T_3", P.T_3); H.tr->append_point("T_4", P.T_4);
        H.tr->append_point("T_5", P.T_5); H.tr->append_point("T_6", P.T_6);
        sc *c = &chal[bi * CH_N];
        sc u = H.tr->challenge_scalar("u"), x = H.tr->challenge_scalar("x");
        c[CH_U] = u; c[CH_X] = x;
        const sc *t = &tout[bi * 8];
        auto horner6 = [&](const sc *v) {   // x * (v0 + x (v1 + ... x v5))
            sc acc = v[5];
            for (int k = 4; k >= 0; k--) acc = sc_add(v[k], sc_mul(x, acc));
            return sc_mul(x, acc);
        };
        P.t_x = horner6(t);
        P.t_x_blinding = horner6(H.tbl);
        P.e_blinding = sc_mul(x, sc_add(H.i_bl, sc_mul(x, sc_add(H.o_bl, sc_mul(x, H.s_bl)))));
        H.tr->append_scalar("t_x", P.t_x);
        H.tr->append_scalar("t_x_blinding", P.t_x_blinding);
        H.tr->append_scalar("e_blinding", P.e_blinding);
        c[CH_W] = H.tr->challenge_scalar("w");
        H.tr->innerproduct_domain_sep(n);
        P.LR.resize((size_t)64 * lg);
    });
    if ((rc = h2d(ctx, ps->chal.p, chal.data(), chal.size() * 32))) return rc;

    // ---- phase 7: inner-product argument, lg n rounds
    k_ipp_init<<<B, BBP_SC_THREADS, 0, ctx->stream>>>(SB);
    ctx->launches++;
    trace.mark("host_ux");
    rc = ipp_rounds(ctx, SB, B, chal, [&](uint32_t j, const std::vector<uint8_t> &lr) {
        parallel_chunks(B, [&](size_t lo, size_t hi) {
            std::vector<sc> inv(hi - lo);
            for (size_t bi = lo; bi < hi; bi++) {
                hstate &H = hs[bi];
                memcpy(&H.pf.LR[(size_t)64 * j], &lr[bi * 64], 64);
                H.tr->append_point("L", &lr[bi * 64]);
                H.tr->append_point("R", &lr[bi * 64 + 32]);
                inv[bi - lo] = chal[bi * CH_N + CH_UJ] = H.tr->challenge_scalar("u");
            }
            sc_batch_invert(inv.data(), inv.size());   // one exponentiation per chunk instead of one per proof
            for (size_t bi = lo; bi < hi; bi++) chal[bi * CH_N + CH_UJINV] = inv[bi - lo];
        });
    }, trace);
    if (rc) return rc;
    std::vector<sc> ab((size_t)B * 2);
    if ((rc = d2h_sync(ctx, ab.data(), ps->ab.p, ab.size() * 32))) return rc;
    for (uint32_t bi = 0; bi < B; bi++) {
        hs[bi].pf.a = ab[bi * 2]; hs[bi].pf.b = ab[bi * 2 + 1];
        S.done(bi, &V[(size_t)bi * m * 32], r1cs_to_bytes(hs[bi].pf, ps->proof_versioned != 0), *hs[bi].tr);
    }
    return 0;
}

// proves jobs[idx[0..B)] — all with the same list length L: Proof::prove (src/blindbid/proof.rs:36-91) over prove_core
inline int prove_group(bbp_ctx *ctx, std::vector<prove_job> &jobs, const std::vector<size_t> &idx) {
    const uint32_t L = (uint32_t)jobs[idx[0]].pub_list.size();
    // the capacity check comes BEFORE the template is built: L is the caller's, the template costs O(L^2) to record
    if (L > BLINDBID_MAX_TOGGLES || next_pow2_u32(blindbid_n1(L)) > ctx->gens_capacity || ctx->party_capacity < 1) {
        for (size_t i : idx) jobs[i].status = BBP_ERR_INVALID_GENERATORS_LENGTH;   // bp_gens.gens_capacity < padded_n
        return 0;
    }
    int rc;
    prove_source S;
    S.B = (uint32_t)idx.size();
    if ((rc = proto_template(ctx, 4, L, &S.dt))) return rc;
    S.inputs = [&](size_t bi, std::vector<sc> &v, std::vector<sc> &bl, const uint8_t *&seed) {
        prove_job &J = jobs[idx[bi]];
        v = {J.d, J.k, J.y, J.y_inv};                                   // src/blindbid/proof.rs:55-58
        for (uint32_t i = 0; i < L; i++) v.push_back(sc_from_u64((uint64_t)i == J.toggle ? 1 : 0));   // proof.rs:60-67
        bl = J.blindings;
        seed = J.rng_seed;
    };
    S.transcript = [](size_t) {
        merlin_transcript *tr = new merlin_transcript("BlindBidProofGadget");   // src/blindbid/mod.rs:37
        tr->r1cs_domain_sep();                                                    // Prover::new
        return tr;
    };
    S.witness = [&](size_t bi, const std::vector<sc> &v, sc *aL, sc *aR, sc *aO) {
        prove_job &J = jobs[idx[bi]];
        std::vector<sc> pub;
        fill_public_values(pub, J.seed, J.q, J.z_img, J.pub_list.data(), L);
        evaluator ev;
        ev.pub = pub.data();
        ev.a_L = aL; ev.a_R = aR; ev.a_O = aO;
        std::vector<sc> toggles(v.begin() + 4, v.end()), items(J.pub_list.begin(), J.pub_list.end());
        proof_gadget(ev, v[0], v[1], v[3], J.q, J.z_img, J.seed, toggles, items);   // proof.rs:74-85
    };
    S.bb_L = L;
    S.bb_witness_in = [&](size_t bi, sc *wi, uint32_t *toggle) {
        prove_job &J = jobs[idx[bi]];
        // a toggle index beyond the list makes every toggle bit zero (the reference: `x as u64 == toggle`); the device
        // witness kernel handles that as well, this only guards the uint32 narrowing
        if (J.toggle > 0xfffffffeull) return false;
        wi[0] = J.d; wi[1] = J.k; wi[2] = J.y_inv; wi[3] = J.seed;
        for (uint32_t i = 0; i < L; i++) wi[4 + i] = J.pub_list[i];
        *toggle = (uint32_t)J.toggle;
        return true;
    };
    S.done = [&](size_t bi, const uint8_t *V, std::vector<uint8_t> &&proof, const merlin_transcript &) {
        prove_job &J = jobs[idx[bi]];
        J.commitments.assign(V, V + 4 * 32);
        J.t_c.assign(V + 4 * 32, V + (size_t)(4 + L) * 32);
        J.proof = std::move(proof);
        J.status = 0;
    };
    S.fail_all = [&](int st) { for (size_t i : idx) jobs[i].status = st; };
    return prove_core(ctx, S);
}

// Lanes: sibling contexts of the same GPU that the 1024-request parts of a larger batch run on concurrently, one host
// thread each (BBP_PROVE_LANES, default 3; 1 = off). No result byte or verdict depends on the cut: requests are
// independent, every per-request draw is keyed by the request's own seed, and a batch verification of several parts is
// the conjunction of one random linear combination per part.
inline size_t batch_lanes() {
    const char *e = getenv("BBP_PROVE_LANES");
    int v = e ? atoi(e) : 3;
    return (size_t)std::min(std::max(v, 1), 8);
}
inline size_t batch_part_max() {
    const char *pe = getenv("BBP_PROVE_PART");   // tests shrink the part size to drive the lanes with a handful of requests
    return pe ? (size_t)std::min(std::max(atoi(pe), 1), 1024) : 1024;
}
// part size for a group of sz requests: two halves for one part of 768..1024 when lanes are available
inline size_t single_part_max(size_t sz) {
    size_t pm = batch_part_max();
    if (!getenv("BBP_PROVE_PART") && batch_lanes() > 1 && sz >= 768 && sz <= 1024) pm = (sz + 1) / 2;
    return pm;
}
// even cuts of at most part_max
inline void cut_parts(const std::vector<size_t> &idx, size_t part_max, std::vector<std::vector<size_t>> &parts) {
    const size_t sz = idx.size(), cuts = (sz + part_max - 1) / part_max, per = (sz + cuts - 1) / cuts;
    for (size_t off = 0; off < sz; off += per) parts.emplace_back(idx.begin() + off, idx.begin() + std::min(sz, off + per));
}

inline int lane_ctx(bbp_ctx *ctx, size_t k, bbp_ctx **out) {
    if (k == 0) { *out = ctx; return 0; }
    while (ctx->lanes.size() < k) {
        bbp_ctx *l = new bbp_ctx();
        l->device = ctx->device;
        int rc = l->init(ctx->gens_capacity, ctx->party_capacity);
        if (rc) { l->destroy(); delete l; return rc; }
        ctx->lanes.push_back(l);
    }
    *out = ctx->lanes[k - 1];
    proto_get(*out)->proof_versioned = proto_get(ctx)->proof_versioned;
    return 0;
}

// fn(lane context, part index) for every part; parts are handed out dynamically to min(lanes, parts) host threads
template <class F>
inline int run_on_lanes(bbp_ctx *ctx, size_t n_parts, F fn) {
    const size_t n_lanes = std::min(batch_lanes(), n_parts);
    if (n_lanes <= 1) {
        for (size_t i = 0; i < n_parts; i++) {
            int rc = fn(ctx, i);
            if (rc) return rc;
        }
        return 0;
    }
    std::vector<bbp_ctx *> lane(n_lanes);
    for (size_t k = 0; k < n_lanes; k++) {
        int rc = lane_ctx(ctx, k, &lane[k]);
        if (rc) return rc;
    }
    std::atomic<size_t> next{0};
    std::vector<int> rcs(n_lanes, 0);
    auto work = [&](size_t k) {
        cudaSetDevice(ctx->device);
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n_parts) return;
            int rc = fn(lane[k], i);
            if (rc) { rcs[k] = rc; return; }
        }
    };
    std::vector<std::thread> th;
    for (size_t k = 1; k < n_lanes; k++) th.emplace_back(work, k);
    work(0);
    for (auto &t : th) t.join();
    for (int rc : rcs) if (rc) return rc;
    return 0;
}

// Proof::prove for a batch of requests; groups by list length
inline int prove_batch(bbp_ctx *ctx, std::vector<prove_job> &jobs) {
    std::map<size_t, std::vector<size_t>> groups;
    for (size_t i = 0; i < jobs.size(); i++) {
        prove_job &J = jobs[i];
        size_t L = J.pub_list.size();
        if (L == 0 || J.blindings.size() != 4 + L) { J.status = BBP_ERR_INPUT; continue; }   // the reference panics (gadgets.rs:103)
        groups[L].push_back(i);
    }
    // parts: at most 1024 proofs each (bounds the device footprint), cut evenly; several parts run on parallel lanes.
    // A single part of 768..1024 is run as two halves on two lanes (105 against 111 ms for 1024); finer cuts, or halving
    // the parts of a larger batch, measured slower (smaller launches lose what the overlap wins).
    std::vector<std::vector<size_t>> parts;
    for (auto &g : groups) cut_parts(g.second, single_part_max(g.second.size()), parts);
    return run_on_lanes(ctx, parts.size(), [&](bbp_ctx *c, size_t i) { return prove_group(c, jobs, parts[i]); });
}


// ================================================================ verifier
struct verify_job {
    // views into the caller's request (valid for the duration of the call): R1CSProof bytes, nc x 32, nt x 32
    const uint8_t *proof = nullptr, *commitments = nullptr, *t_c = nullptr;
    size_t proof_len = 0, n_commitments = 0, n_t_c = 0;
    sc score, z_img, seed;
    std::vector<sc> pub_list;
    uint8_t rng_seed[32];
    int status = 0;    // 0 = accept; BBP_ERR_FORMAT / BBP_ERR_VERIFICATION / BBP_ERR_INVALID_GENERATORS_LENGTH / BBP_ERR_INPUT otherwise
};

struct verify_prepared {
    bool live = false;              // passed every host-side check; takes part in the GPU check
    uint32_t nc = 0, nt = 0, m = 0, n1 = 0, n = 0, lg = 0;
    // [V_0..V_{m-1} | A_I1 A_O1 S1 A_I2 A_O2 S2 | T_1 T_3 T_4 T_5 T_6 | L_0 R_0 .. | t_x t_x_blinding e_blinding | a b], 32 B each:
    // the first m + 11 + 2 lg entries are the dynamic points of the mega-check, the whole blob is what the transcript absorbs
    std::vector<uint8_t> blob;
    std::vector<sc> pub;
    // generic circuits (bbp_r1cs_verify): the caller's template and the transcript state Verifier::new leaves behind
    // (208 B, keccak.h export_state); tr_out, if set, receives the state after the last IPP challenge
    dev_template *dt = nullptr;
    const uint8_t *tr_state = nullptr;
    uint8_t *tr_out = nullptr;
};

// Host part of Verifier::verify for one request: parse (FormatError), the structural checks the reference would panic
// on, the identity checks of validate_and_append_point (VerificationError) and the generator capacity check, in the
// order the reference reports them. The Fiat-Shamir replay itself runs on the device (k_verify_transcript).
inline void verify_prepare(bbp_ctx *ctx, verify_job &J, verify_prepared &P, bool versioned) {
    P.live = false;
    r1cs_proof_host pf;
    if (!r1cs_from_bytes(pf, J.proof, J.proof_len, versioned)) { J.status = BBP_ERR_FORMAT; return; }
    P.nc = (uint32_t)std::min<size_t>(J.n_commitments, 0xffffffffu); P.nt = (uint32_t)std::min<size_t>(J.n_t_c, 0xffffffffu);
    // the reference indexes vars[0], vars[1], vars[3] and toggle[0], items[i] (panics otherwise: SURVEY.md §5)
    if (P.nc < 4 || P.nt < 1 || J.pub_list.size() < P.nt) { J.status = BBP_ERR_FORMAT; return; }
    if (P.nc > BLINDBID_MAX_COMMITMENTS) { J.status = BBP_ERR_INPUT; return; }   // no reference counterpart: bounds the per-request work
    const bool too_many_toggles = P.nt > BLINDBID_MAX_TOGGLES;                   // reported as the capacity error, in its place below
    if (too_many_toggles) P.nt = BLINDBID_MAX_TOGGLES;
    P.m = P.nc + P.nt;
    // n1 by its formula: the template (O(nt^2) to record, cached per (nc, nt)) is only built for requests that pass the
    // generator-capacity check below, i.e. nt <= 202 for BulletproofGens::new(2048, 1)
    P.n1 = blindbid_n1(P.nt); P.n = next_pow2_u32(P.n1); P.lg = log2_u32(P.n);
    if (all_zero32(pf.A_I1) || all_zero32(pf.A_O1) || all_zero32(pf.S1)) { J.status = BBP_ERR_VERIFICATION; return; }
    if (too_many_toggles || P.n > ctx->gens_capacity || ctx->party_capacity < 1) { J.status = BBP_ERR_INVALID_GENERATORS_LENGTH; return; }
    if (all_zero32(pf.T_1) || all_zero32(pf.T_3) || all_zero32(pf.T_4) || all_zero32(pf.T_5) || all_zero32(pf.T_6)) { J.status = BBP_ERR_VERIFICATION; return; }
    // InnerProductProof::verification_scalars
    uint32_t lg_p = (uint32_t)(pf.LR.size() / 64);
    if (lg_p >= 32 || P.n != (1u << lg_p)) { J.status = BBP_ERR_VERIFICATION; return; }
    for (uint32_t j = 0; j < 2 * lg_p; j++)
        if (all_zero32(&pf.LR[32 * (size_t)j])) { J.status = BBP_ERR_VERIFICATION; return; }
    P.blob.resize((size_t)32 * (P.m + 11 + 2 * lg_p + 5));
    uint8_t *o = P.blob.data();
    memcpy(o, J.commitments, (size_t)P.nc * 32); o += (size_t)P.nc * 32;
    memcpy(o, J.t_c, (size_t)P.nt * 32); o += (size_t)P.nt * 32;
    const uint8_t *fixed_pts[11] = {pf.A_I1, pf.A_O1, pf.S1, pf.A_I2, pf.A_O2, pf.S2, pf.T_1, pf.T_3, pf.T_4, pf.T_5, pf.T_6};
    for (int k = 0; k < 11; k++) { memcpy(o, fixed_pts[k], 32); o += 32; }
    memcpy(o, pf.LR.data(), pf.LR.size()); o += pf.LR.size();
    sc_tobytes(o, pf.t_x); sc_tobytes(o + 32, pf.t_x_blinding); sc_tobytes(o + 64, pf.e_blinding); sc_tobytes(o + 96, pf.a); sc_tobytes(o + 128, pf.b);
    // the per-proof part of the public value table (circuit.h: PV_SEED ..): seed, score, z_img, items. The 91 entries before
    // it (1 and the MiMC constants) are the same for every proof and live on the device once (proto_tables)
    P.pub.resize(3 + (size_t)P.nt);
    P.pub[0] = J.seed; P.pub[1] = J.score; P.pub[2] = J.z_img;
    for (uint32_t i = 0; i < P.nt; i++) P.pub[3 + i] = J.pub_list[i];
    P.live = true;
}

// Host twin of k_verify_transcript (used for small batches, where one device thread per request is latency bound):
// fills the request's challenge block (CH_N scalars) and its dynamic scalars [m, ds), unweighted.
inline void verify_transcript_host(const verify_prepared &P, const uint8_t rng_seed[32], sc *c, sc *dyn) {
    const uint32_t m = P.m, lg = P.lg;
    const uint8_t *blob = P.blob.data(), *pts = blob + 32 * (size_t)m, *lr = pts + 32 * 11, *scal = lr + 64 * (size_t)lg;
    merlin_transcript tr("BlindBidProofGadget");
    tr.r1cs_domain_sep();
    if (P.tr_state) tr.import_state(P.tr_state);
    for (uint32_t i = 0; i < m; i++) tr.append_point("V", blob + 32 * (size_t)i);
    tr.append_u64("m", m);
    tr.append_point("A_I1", pts); tr.append_point("A_O1", pts + 32); tr.append_point("S1", pts + 64);
    tr.r1cs_1phase_domain_sep();
    tr.append_point("A_I2", pts + 96); tr.append_point("A_O2", pts + 128); tr.append_point("S2", pts + 160);
    sc y = tr.challenge_scalar("y"), z = tr.challenge_scalar("z");
    tr.append_point("T_1", pts + 192); tr.append_point("T_3", pts + 224); tr.append_point("T_4", pts + 256);
    tr.append_point("T_5", pts + 288); tr.append_point("T_6", pts + 320);
    sc u = tr.challenge_scalar("u"), x = tr.challenge_scalar("x");
    tr.append_message("t_x", scal, 32); tr.append_message("t_x_blinding", scal + 32, 32); tr.append_message("e_blinding", scal + 64, 32);
    sc w = tr.challenge_scalar("w");
    tr.innerproduct_domain_sep(P.n);
    std::vector<sc> inv(lg + 1);
    for (uint32_t j = 0; j < lg; j++) {
        tr.append_point("L", lr + 64 * (size_t)j);
        tr.append_point("R", lr + 64 * (size_t)j + 32);
        inv[j] = c[CH_UJ0 + j] = tr.challenge_scalar("u");
    }
    inv[lg] = y;
    if (P.tr_out) tr.export_state(P.tr_out);
    merlin_rng rng = tr.build_rng().finalize(rng_seed);
    sc r = rng.random_scalar();
    sc_batch_invert(inv.data(), inv.size());
    sc *d = dyn + m;
    for (uint32_t j = 0; j < lg; j++) {
        c[CH_UJ0 + lg + j] = inv[j];
        d[11 + 2 * j] = sc_mul(c[CH_UJ0 + j], c[CH_UJ0 + j]);
        d[11 + 2 * j + 1] = sc_mul(inv[j], inv[j]);
    }
    sc xx = sc_mul(x, x), xxx = sc_mul(xx, x), rxx = sc_mul(r, xx);
    d[0] = x; d[1] = xx; d[2] = xxx; d[3] = sc_mul(u, x); d[4] = sc_mul(u, xx); d[5] = sc_mul(u, xxx);
    d[6] = sc_mul(r, x); d[7] = sc_mul(rxx, x); d[8] = sc_mul(rxx, xx); d[9] = sc_mul(rxx, xxx); d[10] = sc_mul(sc_mul(rxx, xx), xx);
    c[CH_Y] = y; c[CH_YINV] = inv[lg]; c[CH_Z] = z; c[CH_X] = x; c[CH_U] = u; c[CH_W] = w; c[CH_R] = r;
    sc_from_canonical(c[CH_TX], scal); sc_from_canonical(c[CH_TXBL], scal + 32); sc_from_canonical(c[CH_EBL], scal + 64);
    sc_from_canonical(c[CH_A], scal + 96); sc_from_canonical(c[CH_B], scal + 128);
    c[CH_RHO] = sc_one();
}

// GPU part for prepared requests idx[0..B) that share (nc, nt). combined = false: one verdict per request (weight 1);
// combined = true: ONE random linear combination -> a single verdict; the weights come from a Merlin transcript over one
// digest per request — the verifier scalar r that request's own transcript yields after absorbing the whole proof, so it
// binds every proof byte — keyed with batch_seed. Verdicts: 1 = mega-check is the identity.
// Requests with a point that fails to decompress get status VERIFICATION and weight zero.
// A combined pass leaves its per-request state (weighted static rows, weighted dynamic scalars, decompressed points) on the
// context: verify_regroup_pass re-combines it in runs.
inline int verify_group(bbp_ctx *ctx, std::vector<verify_job> &jobs, std::vector<verify_prepared> &prep, const std::vector<size_t> &idx, bool combined,
                        const uint8_t *batch_seed, std::vector<uint8_t> &verdicts, uint8_t *d_partial_ext /* combined: 2 x 128 B, optional */) {
    proto_state *ps = proto_get(ctx);
    const uint32_t B = (uint32_t)idx.size();
    const verify_prepared &P0 = prep[idx[0]];
    int rc;
    if ((rc = proto_tables(ctx))) return rc;
    dev_template *dt = P0.dt;
    if (!dt && (rc = proto_template(ctx, P0.nc, P0.nt, &dt))) return rc;
    const circuit_template &T = *dt->tpl;
    const uint32_t n1 = T.n1, m = T.m, n = P0.n, lg = P0.lg, gcols = ctx->gens_capacity * ctx->party_capacity, slot_len = 2 + 2 * gcols;
    const uint32_t ds = m + 11 + 2 * lg, blob_stride = 32 * (ds + 5);
    const uint32_t gsz = combined ? B : 1;   // requests per verdict
    const uint32_t n_groups = (B + gsz - 1) / gsz;
    phase_trace trace(combined ? "verify_group(combined)" : "verify_group(each)");
    event_timeline tl;

    // ---- upload blobs, seeds, public values; decompress the dynamic points; replay the transcripts
    const uint32_t n_pub_own = T.n_pub - T.n_pub_shared;   // per-proof public values (the shared prefix is resident: ps->pub_shared)
    if ((rc = ps->h_wit.ensure((size_t)B * (blob_stride + 32 + (size_t)n_pub_own * 32) + 32))) return rc;
    uint8_t *h_blobs = ps->h_wit.p, *h_seeds = h_blobs + (size_t)B * blob_stride;   // B request seeds, then the batch seed
    sc *h_pub = (sc *)(h_seeds + (size_t)(B + 1) * 32);
    if (combined) memcpy(h_seeds + (size_t)B * 32, batch_seed, 32); else memset(h_seeds + (size_t)B * 32, 0, 32);
    parallel_for(B, [&](size_t bi) {
        const verify_prepared &P = prep[idx[bi]];
        memcpy(h_blobs + bi * blob_stride, P.blob.data(), blob_stride);
        memcpy(h_seeds + bi * 32, jobs[idx[bi]].rng_seed, 32);
        memcpy(h_pub + bi * n_pub_own, P.pub.data(), (size_t)n_pub_own * 32);
    });
    if ((rc = ps->dyn_pts.ensure((size_t)B * blob_stride)) || (rc = ps->rng_states.ensure((size_t)(B + 1) * 32)) || (rc = ps->pub.ensure((size_t)B * n_pub_own * 32)) ||
        (rc = ps->bw_digests.ensure((size_t)(B / 32 + 1) * 32)) || (rc = ps->dyn_niels.ensure((size_t)B * ds * 96)) || (rc = ps->valid.ensure((size_t)B * ds)) || (rc = ps->chal.ensure((size_t)B * CH_N * 32)) ||
        (rc = ps->dyn_sc.ensure((size_t)B * ds * 32)) || (rc = ps->zpow.ensure((size_t)B * T.q * 32)) || (rc = ps->ypow.ensure((size_t)B * n * 32)) ||
        (rc = ps->yinvpow.ensure((size_t)B * n * 32)) || (rc = ps->stat.ensure((size_t)B * slot_len * 32)) || (rc = ps->sG.ensure((size_t)B * n * 32)) ||
        (rc = ps->stat_red.ensure((size_t)n_groups * slot_len * 32)) || (rc = ps->msm_ext.ensure((size_t)2 * n_groups * 128)) || (rc = ps->flags.ensure(n_groups)))
        return rc;
    trace.mark("stage");
    tl.mark("start", ctx->stream);
    if ((rc = h2d(ctx, ps->dyn_pts.p, h_blobs, (size_t)B * blob_stride)) || (rc = h2d(ctx, ps->rng_states.p, h_seeds, (size_t)(B + 1) * 32)) ||
        (rc = h2d(ctx, ps->pub.p, h_pub, (size_t)B * n_pub_own * 32)))
        return rc;
    // Two streams: the variable-base MSM over the requests' own points (decompression, then the bucket engine) runs on the
    // side stream; transcript replay, weights, power tables, scalar assembly and the static-base MSM run on the main one.
    // The side stream only waits for the weighted dynamic scalars (k_dyn_weights). If the static-base MSM itself needs
    // the bucket engine (per-request checks of a large group, or the latency path switched off) both share the main stream.
    const uint32_t n_stat_slots = n_groups;
    const char *ds_env = getenv("BBP_VERIFY_STREAMS");   // 1 = everything on the main stream (tests compare both)
    const bool two_streams = small_msm_ok((size_t)n_stat_slots * slot_len) && !(ds_env && atoi(ds_env) == 1);
    cudaStream_t side = ctx->stream;
    if (two_streams) {
        if ((rc = proto_side_stream(ps))) return rc;
        side = ps->rng_stream;
        BBP_CUDA_OK(cudaEventRecord(ps->ev_up, ctx->stream));
        BBP_CUDA_OK(cudaStreamWaitEvent(side, ps->ev_up, 0));
    }
    tl.mark("uploaded", ctx->stream);
    k_decompress_to_niels_strided<<<(B * ds + 127) / 128, 128, 0, side>>>(ps->dyn_pts.p, ds, blob_stride, ps->dyn_niels.p, B * ds, ps->valid.p);
    ctx->launches++;
    tl.mark("side:decompressed", side);
    // Fiat-Shamir replay: on the device for large batches (one warp per request, ~0.2 ms whatever the batch), on the host
    // threads otherwise (~40 us per request per thread); BBP_DEVICE_TRANSCRIPT_MIN_BATCH overrides the crossover
    const char *tr_env = getenv("BBP_DEVICE_TRANSCRIPT_MIN_BATCH");
    const bool device_replay = !P0.tr_out && B >= (uint32_t)(tr_env ? atoi(tr_env) : (int)(8 * host_threads()));
    sc_batch SB;
    memset(&SB, 0, sizeof SB);
    SB.n_proofs = B; SB.n1 = n1; SB.q = T.q; SB.m = m; SB.n = n; SB.lg_n = lg; SB.n_pub = T.n_pub; SB.gcols = gcols;
    SB.row_ptr = dt->row_ptr; SB.entries = dt->entries; SB.const_j = dt->const_j; SB.const_idx = dt->const_idx; SB.n_const = (uint32_t)T.const_j.size();
    SB.n_long = dt->n_long;
    memcpy(SB.long_rows, dt->long_rows, sizeof SB.long_rows);
    SB.chal = ps->chal.as<sc>(); SB.zpow = ps->zpow.as<sc>(); SB.ypow = ps->ypow.as<sc>(); SB.yinvpow = ps->yinvpow.as<sc>();
    SB.pub = ps->pub.as<sc>(); SB.dyn_out = ps->dyn_sc.as<sc>(); SB.dyn_stride = ds; SB.stat = ps->stat.as<sc>();
    SB.n_pub = n_pub_own; SB.n_pub_shared = T.n_pub_shared; SB.pub_shared = ps->pub_shared.as<sc>();
    SB.coef = dt->coef;
    SB.stab = ps->sG.as<sc>();
    SB.skip_ypow = 1;
    SB.dyn_done = 1;
    // BBP_VERIFY_SPLIT=1: split replay — phase 1 ends with y, z and y^-1, which is all k_powers needs; the tables are then built
    // on the side stream while phase 2 (a latency chain of ~25 permutations and an inversion) runs here. Measured SLOWER
    // (1024 proofs: 2.6 -> 3.1 ms): the second inversion costs ~0.1 ms and the wide k_powers blocks keep phase 2's warps off
    // the SMs until they drain; kept as an experiment knob, off by default.
    const char *sp_env = getenv("BBP_VERIFY_SPLIT");
    const bool split = two_streams && device_replay && !keccak_per_thread() && sp_env && atoi(sp_env) == 1;
    if ((rc = sc_kernels_smem_opt_in(SB.q, SB.n, SB.lg_n))) return rc;
    if (device_replay) {
        transcript_init init;
        if (P0.tr_state) memcpy(init.state, P0.tr_state, sizeof init.state);   // one group = one circuit = one starting state
        else {
            merlin_transcript tr("BlindBidProofGadget");   // src/blindbid/mod.rs:37
            tr.r1cs_domain_sep();                            // Verifier::new
            tr.export_state(init.state);
        }
        if (keccak_per_thread())
            k_verify_transcript<<<(B + 31) / 32, 32, 0, ctx->stream>>>(init, ps->dyn_pts.p, blob_stride, ps->rng_states.p, B, m, lg, (uint64_t)n,
                                                                       ps->chal.as<sc>(), ps->dyn_sc.as<sc>(), ds);
        else if (!split)   // one warp per request
            k_verify_transcript_warp<<<(B + 3) / 4, 128, 0, ctx->stream>>>(init, ps->dyn_pts.p, blob_stride, ps->rng_states.p, B, m, lg, (uint64_t)n,
                                                                           ps->chal.as<sc>(), ps->dyn_sc.as<sc>(), ds, 0, nullptr);
        else {
            if ((rc = ps->rng_raw.ensure((size_t)B * BBP_STROBE_STATE_BYTES))) return rc;
            k_verify_transcript_warp<<<(B + 3) / 4, 128, 0, ctx->stream>>>(init, ps->dyn_pts.p, blob_stride, ps->rng_states.p, B, m, lg, (uint64_t)n,
                                                                           ps->chal.as<sc>(), ps->dyn_sc.as<sc>(), ds, 1, ps->rng_raw.p);
            BBP_CUDA_OK(cudaEventRecord(ps->ev_head, ctx->stream));
            BBP_CUDA_OK(cudaStreamWaitEvent(side, ps->ev_head, 0));
            k_powers<<<B, BBP_SC_THREADS, k_powers_smem(SB.q, SB.n), side>>>(SB);
            BBP_CUDA_OK(cudaEventRecord(ps->ev_pow, side));
            tl.mark("side:powers", side);
            k_verify_transcript_warp<<<(B + 3) / 4, 128, 0, ctx->stream>>>(init, ps->dyn_pts.p, blob_stride, ps->rng_states.p, B, m, lg, (uint64_t)n,
                                                                           ps->chal.as<sc>(), ps->dyn_sc.as<sc>(), ds, 2, ps->rng_raw.p);
            ctx->launches += 2;
        }
        ctx->launches++;
    } else {
        // pinned staging: the copies below are asynchronous and nothing waits for them before the end of the group
        if ((rc = ps->h_states.ensure(((size_t)B * CH_N + (size_t)B * ds) * 32))) return rc;
        sc *h_chal = (sc *)ps->h_states.p, *h_dyn = h_chal + (size_t)B * CH_N;
        parallel_for(B, [&](size_t bi) {
            for (uint32_t k = 0; k < CH_N; k++) h_chal[bi * CH_N + k] = sc_zero();
            for (uint32_t k = 0; k < ds; k++) h_dyn[bi * ds + k] = sc_zero();
            verify_transcript_host(prep[idx[bi]], jobs[idx[bi]].rng_seed, &h_chal[bi * CH_N], &h_dyn[bi * ds]);
        });
        if ((rc = h2d(ctx, ps->chal.p, h_chal, (size_t)B * CH_N * 32)) || (rc = h2d(ctx, ps->dyn_sc.p, h_dyn, (size_t)B * ds * 32))) return rc;
    }
    trace.mark("h2d+decompress+transcripts(launch)");
    tl.mark("transcripts", ctx->stream);

    // ---- weights, on the device (rng_kernels.cuh): 0 for a request with a point that does not decompress, else 1 (per-request
    // checks) or the request's share of the batch's random linear combination. No host round trip: the whole group is one
    // asynchronous sequence on the stream, synchronised once at the end.
    if (combined) {
        k_batch_weight_digests<<<(B + 31) / 32, 32, 0, ctx->stream>>>(ps->chal.as<sc>(), B, ps->bw_digests.as<uint64_t>());
        ctx->launches++;
    }
    if (two_streams) {   // the validity flags come from the decompression on the side stream
        BBP_CUDA_OK(cudaEventRecord(ps->ev_rng, side));
        BBP_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ps->ev_rng, 0));
    }
    k_batch_weights<<<(B + 127) / 128, 128, 0, ctx->stream>>>(ps->chal.as<sc>(), ps->valid.p, ds, B, ps->bw_digests.as<uint64_t>(),
                                                              ps->rng_states.p + (size_t)B * 32, combined ? 1u : 0u);
    ctx->launches++;

    k_dyn_weights<<<B, 64, 0, ctx->stream>>>(SB);
    ctx->launches++;
    tl.mark("weights", ctx->stream);
    uint8_t *ext = ps->msm_ext.p;
    const uint32_t dyn_total = B * ds, per_slot = gsz * ds;
    {   // dynamic bases: one variable-base slot per group, on the side stream
        if (two_streams) {
            BBP_CUDA_OK(cudaEventRecord(ps->ev_dyn, ctx->stream));
            BBP_CUDA_OK(cudaStreamWaitEvent(side, ps->ev_dyn, 0));
        }
        msm_shape sh = msm_engine::make_shape(dyn_total, per_slot, dyn_total, false, 0, 0, 0);
        cudaStream_t engine_stream = ctx->msm.stream;
        ctx->msm.stream = side;
        rc = ctx->msm.run(sh, ps->dyn_sc.p, ps->dyn_niels.p, ext + (size_t)n_groups * 128, nullptr);
        ctx->msm.stream = engine_stream;
        if (rc) return rc;
        if (two_streams) BBP_CUDA_OK(cudaEventRecord(ps->ev_up, side));
        tl.mark("side:dyn_msm", side);
    }
    if (split) BBP_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ps->ev_pow, 0));
    else k_powers<<<B, BBP_SC_THREADS, k_powers_smem(SB.q, SB.n), ctx->stream>>>(SB);
    tl.mark("powers", ctx->stream);
    k_verify_scalars<<<B, BBP_SC_THREADS, k_verify_scalars_smem(SB.n, SB.lg_n), ctx->stream>>>(SB);
    if (gsz >= 32)
        k_stat_reduce_wide<<<dim3((slot_len + 15) / 16, n_groups), 256, 0, ctx->stream>>>(SB.stat, gsz, slot_len, ps->stat_red.as<sc>(), B);
    else
        k_stat_reduce<<<dim3((slot_len + BBP_SC_THREADS - 1) / BBP_SC_THREADS, n_groups), BBP_SC_THREADS, 0, ctx->stream>>>(SB.stat, gsz, slot_len,
                                                                                                                             ps->stat_red.as<sc>(), B);
    ctx->launches += 3;
    tl.mark("scalars+reduce", ctx->stream);
    // static bases: one fixed-table slot per group
    if ((rc = msm_gens_device(ctx, ps->stat_red.as<sc>(), slot_len, n_groups, nullptr, ext))) return rc;
    tl.mark("static_msm", ctx->stream);
    if (two_streams) BBP_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ps->ev_up, 0));   // join: the dynamic-base sum
    k_group_sum_identity<<<(n_groups + 63) / 64, 64, 0, ctx->stream>>>(ext, n_groups, 2, n_groups, ps->flags.p, nullptr);
    ctx->launches++;
    if (combined && n_groups == 1 && d_partial_ext) BBP_CUDA_OK(cudaMemcpyAsync(d_partial_ext, ext, 256, cudaMemcpyDeviceToDevice, ctx->stream));
    std::vector<uint8_t> fl(n_groups), valid((size_t)B * ds);
    BBP_CUDA_OK(cudaMemcpyAsync(valid.data(), ps->valid.p, valid.size(), cudaMemcpyDeviceToHost, ctx->stream));
    trace.mark("launched");
    tl.mark("end", ctx->stream);
    if ((rc = d2h_sync(ctx, fl.data(), ps->flags.p, n_groups))) return rc;
    trace.mark("sync");
    tl.report(combined ? "verify_group(combined)" : "verify_group(each)");
    ps->resident_B = combined ? B : 0;   // what verify_regroup_pass may re-combine
    ps->resident_ds = ds;
    verdicts.assign(n_groups, 0);
    for (uint32_t g = 0; g < n_groups; g++) verdicts[g] = fl[g];
    for (uint32_t bi = 0; bi < B; bi++) {
        bool alive = true;
        for (uint32_t k = 0; k < ds; k++) alive = alive && valid[(size_t)bi * ds + k];
        if (!alive) jobs[idx[bi]].status = BBP_ERR_VERIFICATION;   // optional_multiscalar_mul -> None
        else if (!combined) jobs[idx[bi]].status = fl[bi] ? 0 : BBP_ERR_VERIFICATION;
    }
    return 0;
}

// Narrowing pass after a failed combination: the B requests the last combined verify_group call on THIS context checked are
// re-combined in consecutive runs of g (the same weights; ceil(B / g) verdicts) from the state that call left in HBM — the
// weighted static rows (ps->stat), the weighted dynamic scalars and the decompressed points — so it costs two small MSMs
// and no upload, replay or scalar assembly.
inline int verify_regroup_pass(bbp_ctx *ctx, uint32_t B, uint32_t g, std::vector<uint8_t> &verdicts) {
    proto_state *ps = proto_get(ctx);
    if (!B || !g || ps->resident_B != B) return BBP_ERR_INPUT;
    const uint32_t ds = ps->resident_ds, gcols = ctx->gens_capacity * ctx->party_capacity, slot_len = 2 + 2 * gcols, n_groups = (B + g - 1) / g;
    int rc;
    if ((rc = ps->stat_red.ensure((size_t)n_groups * slot_len * 32)) || (rc = ps->msm_ext.ensure((size_t)2 * n_groups * 128)) || (rc = ps->flags.ensure(n_groups)))
        return rc;
    if (g >= 32)
        k_stat_reduce_wide<<<dim3((slot_len + 15) / 16, n_groups), 256, 0, ctx->stream>>>(ps->stat.as<sc>(), g, slot_len, ps->stat_red.as<sc>(), B);
    else
        k_stat_reduce<<<dim3((slot_len + BBP_SC_THREADS - 1) / BBP_SC_THREADS, n_groups), BBP_SC_THREADS, 0, ctx->stream>>>(ps->stat.as<sc>(), g, slot_len,
                                                                                                                             ps->stat_red.as<sc>(), B);
    ctx->launches++;
    uint8_t *ext = ps->msm_ext.p;
    if ((rc = msm_gens_device(ctx, ps->stat_red.as<sc>(), slot_len, n_groups, nullptr, ext))) return rc;
    msm_shape sh = msm_engine::make_shape(B * ds, g * ds, B * ds, false, 0, 0, 0);
    if ((rc = ctx->msm.run(sh, ps->dyn_sc.p, ps->dyn_niels.p, ext + (size_t)n_groups * 128, nullptr))) return rc;
    k_group_sum_identity<<<(n_groups + 63) / 64, 64, 0, ctx->stream>>>(ext, n_groups, 2, n_groups, ps->flags.p, nullptr);
    ctx->launches++;
    verdicts.assign(n_groups, 0);
    ps->resident_B = 0;   // stat_red / msm_ext were reused; the per-request state itself is still intact but one narrowing pass is all there is
    return d2h_sync(ctx, verdicts.data(), ps->flags.p, n_groups);
}

inline void verify_prepare_all(bbp_ctx *ctx, std::vector<verify_job> &jobs, std::vector<verify_prepared> &prep, std::map<uint64_t, std::vector<size_t>> &groups) {
    proto_state *ps = proto_get(ctx);
    bool versioned = ps->proof_versioned != 0;
    {
        phase_trace tp("verify_prepare(host)");
        parallel_for(jobs.size(), [&](size_t i) { verify_prepare(ctx, jobs[i], prep[i], versioned); });
        tp.mark("parse");
    }
    for (size_t i = 0; i < jobs.size(); i++)
        if (prep[i].live) groups[((uint64_t)prep[i].nc << 32) | prep[i].nt].push_back(i);
}

// Verify::verify for every request independently (exactly the reference's per-request semantics)
inline int verify_each(bbp_ctx *ctx, std::vector<verify_job> &jobs) {
    std::vector<verify_prepared> prep(jobs.size());
    std::map<uint64_t, std::vector<size_t>> groups;
    verify_prepare_all(ctx, jobs, prep, groups);
    std::vector<std::vector<size_t>> parts;
    for (auto &g : groups) cut_parts(g.second, batch_part_max(), parts);
    return run_on_lanes(ctx, parts.size(), [&](bbp_ctx *c, size_t i) {
        std::vector<uint8_t> verdicts;
        return verify_group(c, jobs, prep, parts[i], false, nullptr, verdicts, nullptr);
    });
}

// run length of the narrowing pass after a failed combination (BBP_VERIFY_REGROUP, 0 / 1 = off: per-request pass at once)
inline uint32_t verify_regroup() {
    const char *e = getenv("BBP_VERIFY_REGROUP");
    int v = e ? atoi(e) : 8;
    return v < 0 ? 0u : (uint32_t)v;
}

// Batch verification: one random linear combination of the mega-checks of all requests of a
// circuit shape. If the combination is the identity every live request is accepted; otherwise the requests are
// re-checked individually, so the verdicts always equal those of verify_each.
// partial_only: stop after the combined pass and leave this GPU's partial sum (static | dynamic, 2 x 128 B extended) in
// d_partial_ext for a cross-GPU reduction (proof-range sharding); *all_ok then reports the local verdict:
// every local request passed the host checks and decompressed AND the local combination is the identity. The whole batch
// verifies iff every rank's local flag is set and the sum of all partials is the identity.
inline int verify_batch(bbp_ctx *ctx, std::vector<verify_job> &jobs, const uint8_t batch_seed[32], int *all_ok, bool partial_only, uint8_t *d_partial_ext) {
    std::vector<verify_prepared> prep(jobs.size());
    std::map<uint64_t, std::vector<size_t>> groups;
    verify_prepare_all(ctx, jobs, prep, groups);
    if (partial_only && groups.size() > 1) return BBP_ERR_INPUT;   // sharded mode expects one circuit shape per call
    // one combination per part of at most 1024 requests; the parts run on parallel lanes. In the sharded mode every part
    // leaves its partial sum (static | dynamic, 2 x 128 B) in a scratch row and the rows are folded into the one partial
    // this GPU contributes to the cross-GPU sum.
    std::vector<std::vector<size_t>> parts;
    for (auto &g : groups) cut_parts(g.second, single_part_max(g.second.size()), parts);
    proto_state *ps0 = proto_get(ctx);
    if (partial_only && !parts.empty()) {
        int r = ps0->lane_partials.ensure(parts.size() * 256);
        if (r) return r;
    }
    std::vector<uint8_t> part_ok(parts.size(), 0);
    int rc = run_on_lanes(ctx, parts.size(), [&](bbp_ctx *c, size_t i) {
        std::vector<uint8_t> verdicts;
        int r = verify_group(c, jobs, prep, parts[i], true, batch_seed, verdicts, partial_only ? ps0->lane_partials.p + 256 * i : nullptr);
        if (r) return r;
        part_ok[i] = verdicts[0];
        if (!verdicts[0] && !partial_only) {
            // The combination failed: find the culprits. A per-request pass over the whole part costs ~12 us per request
            // (4098 static columns each); with few culprits it is cheaper to re-combine runs of g requests first (same
            // weights, ceil(B / g) verdicts, from the state the combined pass left on this lane) and to check one by one only
            // the runs that fail.
            const std::vector<size_t> *suspects = &parts[i];
            std::vector<size_t> narrowed;
            const uint32_t g = verify_regroup();
            if (g >= 2 && parts[i].size() >= 4 * (size_t)g) {
                std::vector<uint8_t> vg;
                if ((r = verify_regroup_pass(c, (uint32_t)parts[i].size(), g, vg))) return r;
                for (size_t k = 0; k < parts[i].size(); k++)
                    if (!vg[k / g]) narrowed.push_back(parts[i][k]);
                // every run passing while their sum fails cannot happen; if it does, trust nothing and check everyone
                if (!narrowed.empty()) suspects = &narrowed;
            }
            for (size_t k : *suspects) jobs[k].status = 0;
            std::vector<uint8_t> v2;
            if ((r = verify_group(c, jobs, prep, *suspects, false, nullptr, v2, nullptr))) return r;
        }
        return 0;
    });
    if (!rc && partial_only && !parts.empty()) {
        // every lane has synchronised its own stream inside verify_group, so the rows are complete
        k_partial_fold<<<1, 32, 0, ctx->stream>>>(ps0->lane_partials.p, (uint32_t)parts.size(), 2, d_partial_ext);
        ctx->launches++;
        { int wr = wait_stream(ctx); if (wr) return wr; }
    }
    if (!rc && partial_only && parts.empty()) {   // no live request: this GPU contributes the identity twice
        uint8_t id[256];
        memset(id, 0, sizeof id);
        id[32] = 1; id[64] = 1; id[128 + 32] = 1; id[128 + 64] = 1;   // (X, Y, Z, T) = (0, 1, 1, 0)
        BBP_CUDA_OK(cudaMemcpyAsync(d_partial_ext, id, 256, cudaMemcpyHostToDevice, ctx->stream));
        { int wr = wait_stream(ctx); if (wr) return wr; }
    }
    if (rc) return rc;
    bool ok = true;
    for (uint8_t v : part_ok) ok = ok && v;
    // a request that failed a host check (format, identity point, non-canonical scalar, wrong IPP length) or whose points do
    // not decompress never entered a combination: it must fail the batch in BOTH modes (in the sharded mode the caller
    // ANDs this flag over the ranks next to the identity test of the summed partials)
    for (auto &J : jobs) ok = ok && J.status == 0;
    if (all_ok) *all_ok = ok ? 1 : 0;
    return 0;
}


// ================================================================ generic bulletproofs surface (layer 3)
// What src/gadgets.rs and src/blindbid compile against: ConstraintSystem::{multiply, constrain} (gadgets.rs:53, 30),
// Prover::{new, commit, prove} (blindbid/proof.rs:50, 57, 88), Verifier::{new, commit, verify} (blindbid/verify.rs:51, 57,
// 88) and InnerProductProof::create underneath. The caller records its circuit (and, proving, its assignments) and hands
// the flattened form over; the prover / verifier below are the same ones the blind-bid entry points run on.
struct generic_cs {
    std::shared_ptr<const circuit_template> tpl;
};
inline int generic_cs_build(generic_cs &out, uint32_t n_mul, uint32_t n_commit, uint32_t n_con, const uint32_t *con_ptr, const uint32_t *term_var,
                            const uint8_t *term_coeff) {
    // the power tables address z^(j+1), j < q, and y^i, i < n, through 16-bit two-level indices (sc_kernels.cuh: k_powers,
    // k_dyn_weights): q + 1 and the padded n must not exceed 65536
    if (n_mul == 0 || n_mul > (1u << 16) || n_commit > (1u << 16) || n_con >= (1u << 16)) return BBP_ERR_INPUT;
    out.tpl = generic_template(n_mul, n_commit, n_con, con_ptr, term_var, term_coeff);
    return out.tpl ? 0 : BBP_ERR_FORMAT;
}

// Prover::prove over a recorded circuit. tr: the caller's transcript as Transcript::new(label) left it (Prover::new's
// domain separator is applied here); on success it is advanced exactly as the Rust prover advances its &mut Transcript.
inline int r1cs_prove_generic(bbp_ctx *ctx, merlin_transcript &tr, const generic_cs &cs, const sc *aL, const sc *aR, const sc *aO, const sc *v,
                              const sc *v_blinding, const uint8_t rng_seed[32], uint8_t *V_out, std::vector<uint8_t> &proof, int *status) {
    const circuit_template &T = *cs.tpl;
    dev_template dt;
    dt.tpl = cs.tpl;
    int rc = template_upload(ctx, dt);
    if (rc) { template_free(dt); return rc; }
    prove_source S;
    S.B = 1;
    S.dt = &dt;
    S.inputs = [&](size_t, std::vector<sc> &vv, std::vector<sc> &bl, const uint8_t *&seed) {
        vv.assign(v, v + T.m); bl.assign(v_blinding, v_blinding + T.m); seed = rng_seed;
    };
    S.transcript = [&](size_t) {
        merlin_transcript *t = new merlin_transcript(tr);
        t->r1cs_domain_sep();   // Prover::new
        return t;
    };
    S.witness = [&](size_t, const std::vector<sc> &, sc *l, sc *r, sc *o) {
        memcpy(l, aL, (size_t)T.n1 * 32); memcpy(r, aR, (size_t)T.n1 * 32); memcpy(o, aO, (size_t)T.n1 * 32);
    };
    *status = BBP_ERR_CUDA;
    S.done = [&](size_t, const uint8_t *V, std::vector<uint8_t> &&pf, const merlin_transcript &after) {
        if (V_out) memcpy(V_out, V, (size_t)T.m * 32);
        proof = std::move(pf);
        tr = after;
        *status = 0;
    };
    S.fail_all = [&](int st) { *status = st; };
    rc = prove_core(ctx, S);
    cudaStreamSynchronize(ctx->stream);
    template_free(dt);
    return rc;
}

// Verifier::verify over a recorded circuit: 0 = accept, BBP_ERR_* mirroring R1CSError otherwise
inline int r1cs_verify_generic(bbp_ctx *ctx, merlin_transcript &tr, const generic_cs &cs, const uint8_t *proof, size_t proof_len, const uint8_t *V,
                               const uint8_t rng_seed[32], int *status) {
    const circuit_template &T = *cs.tpl;
    proto_state *ps = proto_get(ctx);
    std::vector<verify_job> jobs(1);
    std::vector<verify_prepared> prep(1);
    verify_job &J = jobs[0];
    verify_prepared &P = prep[0];
    memcpy(J.rng_seed, rng_seed, 32);
    J.status = 0;
    *status = 0;
    r1cs_proof_host pf;
    if (!r1cs_from_bytes(pf, proof, proof_len, ps->proof_versioned != 0)) { *status = BBP_ERR_FORMAT; return 0; }
    P.m = T.m; P.n1 = T.n1; P.n = next_pow2_u32(T.n1); P.lg = log2_u32(P.n);
    // the order of the checks is Verifier::verify's (see verify_prepare)
    if (all_zero32(pf.A_I1) || all_zero32(pf.A_O1) || all_zero32(pf.S1)) { *status = BBP_ERR_VERIFICATION; return 0; }
    if (P.n > ctx->gens_capacity || ctx->party_capacity < 1) { *status = BBP_ERR_INVALID_GENERATORS_LENGTH; return 0; }
    if (all_zero32(pf.T_1) || all_zero32(pf.T_3) || all_zero32(pf.T_4) || all_zero32(pf.T_5) || all_zero32(pf.T_6)) { *status = BBP_ERR_VERIFICATION; return 0; }
    const uint32_t lg_p = (uint32_t)(pf.LR.size() / 64);
    if (lg_p >= 32 || P.n != (1u << lg_p)) { *status = BBP_ERR_VERIFICATION; return 0; }
    for (uint32_t j = 0; j < 2 * lg_p; j++)
        if (all_zero32(&pf.LR[32 * (size_t)j])) { *status = BBP_ERR_VERIFICATION; return 0; }
    P.blob.resize((size_t)32 * (P.m + 11 + 2 * lg_p + 5));
    uint8_t *o = P.blob.data();
    memcpy(o, V, (size_t)P.m * 32); o += (size_t)P.m * 32;
    const uint8_t *fixed_pts[11] = {pf.A_I1, pf.A_O1, pf.S1, pf.A_I2, pf.A_O2, pf.S2, pf.T_1, pf.T_3, pf.T_4, pf.T_5, pf.T_6};
    for (int k = 0; k < 11; k++) { memcpy(o, fixed_pts[k], 32); o += 32; }
    memcpy(o, pf.LR.data(), pf.LR.size()); o += pf.LR.size();
    sc_tobytes(o, pf.t_x); sc_tobytes(o + 32, pf.t_x_blinding); sc_tobytes(o + 64, pf.e_blinding); sc_tobytes(o + 96, pf.a); sc_tobytes(o + 128, pf.b);
    P.pub = T.coef_table;
    P.live = true;
    dev_template dt;
    dt.tpl = cs.tpl;
    int rc = template_upload(ctx, dt);
    if (rc) { template_free(dt); return rc; }
    uint8_t st_in[BBP_STROBE_STATE_BYTES], st_out[BBP_STROBE_STATE_BYTES];
    {
        merlin_transcript t0(tr);
        t0.r1cs_domain_sep();   // Verifier::new
        t0.export_state(st_in);
    }
    P.dt = &dt; P.tr_state = st_in; P.tr_out = st_out;
    std::vector<size_t> idx(1, 0);
    std::vector<uint8_t> verdicts;
    rc = verify_group(ctx, jobs, prep, idx, false, nullptr, verdicts, nullptr);
    template_free(dt);
    if (rc) return rc;
    tr.import_state(st_out);
    *status = J.status;
    return 0;
}

// InnerProductProof::create with Q = w * B (B = the Pedersen value base: how Prover::prove and RangeProof use it), over
// the first n resident generators G, H of party 0. Output: L_0 R_0 .. L_{lg n - 1} R_{lg n - 1} a b.
inline int ipp_create_generic(bbp_ctx *ctx, merlin_transcript &tr, const sc &w, const sc *Gf, const sc *Hf, const sc *a, const sc *b, uint32_t n,
                              std::vector<uint8_t> &out) {
    proto_state *ps = proto_get(ctx);
    int rc;
    if (n == 0 || (n & (n - 1)) || n > ctx->gens_capacity || ctx->party_capacity < 1) return BBP_ERR_INVALID_GENERATORS_LENGTH;
    if ((rc = proto_tables(ctx))) return rc;
    const uint32_t lg = log2_u32(n), gcols = ctx->gens_capacity * ctx->party_capacity, slot_len = 2 + 2 * gcols;
    if ((rc = ps->chal.ensure((size_t)CH_N * 32)) || (rc = ps->a.ensure((size_t)n * 32)) || (rc = ps->b.ensure((size_t)n * 32)) ||
        (rc = ps->sG.ensure((size_t)n * 32)) || (rc = ps->sH.ensure((size_t)n * 32)) || (rc = ps->slots.ensure((size_t)3 * slot_len * 32)) ||
        (rc = ps->ab.ensure(64)) || (rc = ps->wit.ensure((size_t)4 * n * 32)))
        return rc;
    sc_batch SB;
    memset(&SB, 0, sizeof SB);
    SB.n_proofs = 1; SB.n1 = n; SB.n = n; SB.lg_n = lg; SB.gcols = gcols;
    SB.chal = ps->chal.as<sc>(); SB.a = ps->a.as<sc>(); SB.b = ps->b.as<sc>(); SB.sG = ps->sG.as<sc>(); SB.sH = ps->sH.as<sc>();
    SB.slots = ps->slots.as<sc>(); SB.ab_out = ps->ab.as<sc>();
    sc *stage = ps->wit.as<sc>();
    if ((rc = h2d(ctx, stage, a, (size_t)n * 32)) || (rc = h2d(ctx, stage + n, b, (size_t)n * 32)) || (rc = h2d(ctx, stage + 2 * (size_t)n, Gf, (size_t)n * 32)) ||
        (rc = h2d(ctx, stage + 3 * (size_t)n, Hf, (size_t)n * 32)))
        return rc;
    std::vector<sc> chal(CH_N, sc_zero());
    chal[CH_W] = w;
    if ((rc = h2d(ctx, ps->chal.p, chal.data(), chal.size() * 32))) return rc;
    { int wr = wait_stream(ctx); if (wr) return wr; }   // the sources are pageable caller memory
    k_ipp_load<<<1, BBP_SC_THREADS, 0, ctx->stream>>>(SB, stage, stage + n, stage + 2 * (size_t)n, stage + 3 * (size_t)n);
    ctx->launches++;
    tr.innerproduct_domain_sep(n);
    out.assign((size_t)64 * lg + 64, 0);
    phase_trace trace("ipp_create");
    rc = ipp_rounds(ctx, SB, 1, chal, [&](uint32_t j, const std::vector<uint8_t> &lr) {
        memcpy(&out[(size_t)64 * j], lr.data(), 64);
        tr.append_point("L", lr.data());
        tr.append_point("R", lr.data() + 32);
        sc u = tr.challenge_scalar("u");
        chal[CH_UJ] = u;
        chal[CH_UJINV] = sc_invert(u);
    }, trace);
    if (rc) return rc;
    sc ab[2];
    if ((rc = d2h_sync(ctx, ab, ps->ab.p, 64))) return rc;
    sc_tobytes(&out[(size_t)64 * lg], ab[0]);
    sc_tobytes(&out[(size_t)64 * lg + 32], ab[1]);
    return 0;
}

}  // namespace bbp
