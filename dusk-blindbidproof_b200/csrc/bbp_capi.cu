// C ABI of the B200 backend (include/bbp.h): context, resident generator tables, base tables, MSM entry points,
// point codecs and the primitive test hooks. The protocol layer (R1CS prove / verify, blind-bid drivers) lives in
// r1cs.cuh / blindbid.cuh and is exported from the bottom of this translation unit.
#include <sys/random.h>
#include <cerrno>
#include "../../include/bbp.h"
#include <time.h>
#include "codec.cuh"
#include "ctx.cuh"
#include "msm.cuh"
#include "protocol.cuh"
#include "wire.h"
#include "rangeproof.cuh"

using namespace bbp;

extern "C" {

int bbp_init(bbp_ctx **out, int device, uint32_t gens_capacity, uint32_t party_capacity) {
    if (!out) return BBP_ERR_INPUT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        fprintf(stderr, "bbp_init: no usable CUDA device %d (found %d); this backend has no CPU fallback\n", device, count);
        return BBP_ERR_CUDA;
    }
    bbp_ctx *ctx = new bbp_ctx();
    ctx->device = device;
    int rc = ctx->init(gens_capacity, party_capacity);
    if (rc) { delete ctx; return rc; }
    *out = ctx;
    return BBP_OK;
}

void bbp_free(bbp_ctx *ctx) {
    if (!ctx) return;
    ctx->destroy();
    delete ctx;
}

uint64_t bbp_launch_count(const bbp_ctx *ctx) {
    if (!ctx) return 0;
    uint64_t n = ctx->launches + ctx->msm.launches;
    for (const bbp_ctx *l : ctx->lanes) n += l->launches + l->msm.launches;
    return n;
}
uint64_t bbp_stream(const bbp_ctx *ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }
int bbp_sync(bbp_ctx *ctx) {
    if (!ctx) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (bbp_ctx *l : ctx->lanes) BBP_CUDA_OK(cudaStreamSynchronize(l->stream));
    return BBP_OK;
}

int bbp_lane(bbp_ctx *ctx, uint32_t k, bbp_ctx **lane) {
    if (!ctx || !lane || k > 7) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    return lane_ctx(ctx, k, lane);
}

int bbp_pedersen_gens(bbp_ctx *ctx, uint8_t B[32], uint8_t B_blinding[32]) {
    if (!ctx || !B || !B_blinding) return BBP_ERR_INPUT;
    memcpy(B, ctx->pc_compressed, 32);
    memcpy(B_blinding, ctx->pc_compressed + 32, 32);
    return BBP_OK;
}

int bbp_bulletproof_gens(bbp_ctx *ctx, int which, uint32_t party, uint32_t first, uint32_t count, uint8_t *out) {
    if (!ctx || !out || (which != 'G' && which != 'H')) return BBP_ERR_INPUT;
    if (party >= ctx->party_capacity || (uint64_t)first + count > ctx->gens_capacity) return BBP_ERR_INVALID_GENERATORS_LENGTH;
    cudaSetDevice(ctx->device);
    size_t idx = ctx->gen_index(which, party, first);
    uint8_t *d_tmp = nullptr;
    BBP_CUDA_OK(cudaMalloc(&d_tmp, (size_t)count * 32));
    k_compress<<<(count + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_gens_ext + 128 * idx, (uint32_t *)d_tmp, count);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out, d_tmp, (size_t)count * 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
    return e == cudaSuccess ? BBP_OK : BBP_ERR_CUDA;
}

// ---------------------------------------------------------------- measurement hooks
int bbp_set_profiling(bbp_ctx *ctx, int on) {
    if (!ctx) return BBP_ERR_INPUT;
    ctx->msm.profile = on != 0;
    return BBP_OK;
}
int bbp_msm_stage_ms(bbp_ctx *ctx, float *ms, size_t n_stages) {
    if (!ctx || !ms || n_stages != (size_t)msm_engine::N_STAGES) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    return ctx->msm.collect(ms) ? BBP_ERR_CUDA : BBP_OK;
}
int bbp_msm_plan(size_t n, uint32_t out[4]) {
    if (!out || n == 0 || n > 0x7fffffffu) return BBP_ERR_INPUT;
    msm_shape sh = msm_engine::make_shape((uint32_t)n, (uint32_t)n, (uint32_t)n, false, 0, 0, 0);
    out[0] = sh.c; out[1] = sh.W; out[2] = sh.S; out[3] = sh.B;
    return BBP_OK;
}
static int int_peak_run(bbp_ctx *ctx, int pairs, double *per_s, double *per_clk_per_sm) {
    if (!ctx || !per_s) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    cudaDeviceProp prop;
    BBP_CUDA_OK(cudaGetDeviceProperties(&prop, ctx->device));
    int blocks = prop.multiProcessorCount * 2, threads = 1024;
    uint32_t *d = nullptr;
    unsigned long long *d_cyc = nullptr;
    BBP_CUDA_OK(cudaMalloc(&d, (size_t)blocks * threads * 4));
    BBP_CUDA_OK(cudaMalloc(&d_cyc, (size_t)blocks * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    std::vector<unsigned long long> cyc(blocks);
    double best_per_clk = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0, ctx->stream);
        if (pairs) k_int_peak_pair<<<blocks, threads, 0, ctx->stream>>>(d, rep + 1, d_cyc);
        else k_int_peak<<<blocks, threads, 0, ctx->stream>>>(d, rep + 1, d_cyc);
        cudaEventRecord(e1, ctx->stream);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); cudaFree(d_cyc); return BBP_ERR_CUDA; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
        cudaMemcpy(cyc.data(), d_cyc, (size_t)blocks * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < blocks; i++) avg += (double)cyc[i];
        avg /= blocks;
        // two resident CTAs per SM share the multiplier for `avg` SM clocks
        double per_clk = 2.0 * threads * (double)BBP_PEAK_ITERS * BBP_PEAK_ILP / avg;
        if (rep && per_clk > best_per_clk) best_per_clk = per_clk;
    }
    ctx->launches += 6;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d); cudaFree(d_cyc);
    *per_s = (double)blocks * threads * (double)BBP_PEAK_ITERS * BBP_PEAK_ILP / (best * 1e-3);
    if (per_clk_per_sm) *per_clk_per_sm = best_per_clk;
    return BBP_OK;
}
int bbp_int_peak(bbp_ctx *ctx, double *wide_mads_per_s, double *wide_mads_per_clk_per_sm) { return int_peak_run(ctx, 0, wide_mads_per_s, wide_mads_per_clk_per_sm); }
int bbp_int_peak_pairs(bbp_ctx *ctx, double *pairs_per_s, double *pairs_per_clk_per_sm) { return int_peak_run(ctx, 1, pairs_per_s, pairs_per_clk_per_sm); }

// ---------------------------------------------------------------- base tables
int bbp_points_from_compressed(bbp_ctx *ctx, const uint8_t *points, size_t n, bbp_points **out, int *all_valid) {
    if (!ctx || !points || !out || n == 0 || n > 0x7fffffffu) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    bbp_points *p = new bbp_points();
    p->ctx = ctx; p->n = n;
    uint8_t *d_in = nullptr;
    int *d_valid = nullptr;
    int rc = BBP_OK, h_valid = 1;
    if (cudaMalloc(&p->d_niels, n * 96) != cudaSuccess || cudaMalloc(&d_in, n * 32) != cudaSuccess || cudaMalloc(&d_valid, 4) != cudaSuccess) rc = BBP_ERR_CUDA;
    if (!rc && cudaMemcpyAsync(d_in, points, n * 32, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = BBP_ERR_CUDA;
    if (!rc && cudaMemcpyAsync(d_valid, &h_valid, 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = BBP_ERR_CUDA;
    if (!rc) {
        k_decompress_to_niels<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t *)d_in, p->d_niels, (uint32_t)n, d_valid, nullptr);
        ctx->launches++;
        if (cudaMemcpyAsync(&h_valid, d_valid, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = BBP_ERR_CUDA;
    }
    cudaFree(d_in); cudaFree(d_valid);
    if (rc) { cudaFree(p->d_niels); delete p; return rc; }
    if (all_valid) *all_valid = h_valid;
    *out = p;
    return BBP_OK;
}

int bbp_points_from_extended(bbp_ctx *ctx, const uint8_t *points_ext, size_t n, bbp_points **out) {
    if (!ctx || !points_ext || !out || n == 0 || n > 0x7fffffffu) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    bbp_points *p = new bbp_points();
    p->ctx = ctx; p->n = n;
    uint8_t *d_in = nullptr;
    int rc = BBP_OK;
    if (cudaMalloc(&p->d_niels, n * 96) != cudaSuccess || cudaMalloc(&d_in, n * 128) != cudaSuccess) rc = BBP_ERR_CUDA;
    if (!rc && cudaMemcpyAsync(d_in, points_ext, n * 128, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = BBP_ERR_CUDA;
    if (!rc) {
        size_t threads = (n + BBP_NIELS_BATCH - 1) / BBP_NIELS_BATCH;
        k_ext_to_niels<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(d_in, p->d_niels, (uint32_t)n);
        ctx->launches++;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = BBP_ERR_CUDA;
    }
    cudaFree(d_in);
    if (rc) { cudaFree(p->d_niels); delete p; return rc; }
    *out = p;
    return BBP_OK;
}

size_t bbp_points_len(const bbp_points *p) { return p ? p->n : 0; }
void bbp_points_free(bbp_points *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaFree(p->d_niels);
    delete p;
}

// ---------------------------------------------------------------- MSM
int bbp_msm_points_device(bbp_ctx *ctx, const void *scalars_device, size_t n, const bbp_points *points, void *out_device, void *out_ext_device) {
    if (!ctx || !scalars_device || !points || (!out_device && !out_ext_device) || n == 0 || n != points->n) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    msm_shape sh = msm_engine::make_shape((uint32_t)n, (uint32_t)n, (uint32_t)n, false, 0, 0, 0);
    return ctx->msm.run(sh, (const uint8_t *)scalars_device, points->d_niels, (uint8_t *)out_ext_device, (uint8_t *)out_device);
}

int bbp_sum_compress_device(bbp_ctx *ctx, const void *points_ext_device, size_t n, void *out_device) {
    if (!ctx || !points_ext_device || !out_device || n == 0 || n > 1024) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    k_sum_compress<<<1, 32, 0, ctx->stream>>>((const uint8_t *)points_ext_device, (uint32_t)n, (uint32_t *)out_device);
    ctx->launches++;
    BBP_CUDA_OK(cudaGetLastError());
    return BBP_OK;
}

int bbp_sharded_verdict_device(bbp_ctx *ctx, const void *rows_device, size_t world, size_t row_stride, void *out_device) {
    if (!ctx || !rows_device || !out_device || world == 0 || world > 4096 || row_stride < 257) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    k_sharded_verdict<<<1, 32, 0, ctx->stream>>>((const uint8_t *)rows_device, (uint32_t)world, (uint32_t)row_stride, (uint8_t *)out_device);
    ctx->launches++;
    BBP_CUDA_OK(cudaGetLastError());
    return BBP_OK;
}

int bbp_msm_points_batched(bbp_ctx *ctx, const uint8_t *scalars, size_t n_per_slot, size_t n_slots, const bbp_points *points, uint8_t *out) {
    if (!ctx || !scalars || !points || !out || n_per_slot == 0 || n_slots == 0 || n_per_slot != points->n) return BBP_ERR_INPUT;
    size_t n = n_per_slot * n_slots;
    if (n > 0x7fffffffu) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    int rc = ctx->stage_in(scalars, n * 32);
    if (rc) return rc;
    rc = ctx->reserve_out(n_slots * 32);
    if (rc) return rc;
    msm_shape sh = msm_engine::make_shape((uint32_t)n, (uint32_t)n_per_slot, (uint32_t)n_per_slot, false, 0, 0, 0);
    rc = ctx->msm.run(sh, ctx->d_in, points->d_niels, nullptr, ctx->d_out);
    if (rc) return rc;
    BBP_CUDA_OK(cudaMemcpyAsync(out, ctx->d_out, n_slots * 32, cudaMemcpyDeviceToHost, ctx->stream));
    BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BBP_OK;
}

int bbp_msm_points(bbp_ctx *ctx, const uint8_t *scalars, size_t n, const bbp_points *points, uint8_t out[32]) {
    return bbp_msm_points_batched(ctx, scalars, n, 1, points, out);
}

// one-shot forms: temporary base table and staging live in grow-only context scratch (no allocation per call)
static int msm_oneshot(bbp_ctx *ctx, const uint8_t *scalars, const uint8_t *points, size_t n, bool compressed, uint8_t out[32]) {
    if (ctx && out && n == 0) { memset(out, 0, 32); return BBP_OK; }   // the empty sum is the identity, as in dalek
    if (!ctx || !scalars || !points || !out || n > 0x7fffffffu) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    phase_trace trace("msm_oneshot");
    const size_t pt_bytes = compressed ? 32 : 128;
    int rc;
    if ((rc = ctx->reserve_scratch(n * 96 + n * pt_bytes + 16))) return rc;
    uint8_t *d_table = ctx->d_scratch, *d_pts = ctx->d_scratch + n * 96;
    int *d_valid = (int *)(d_pts + n * pt_bytes);
    int h_valid = 1;
    // scalars go first on the main stream and their recode / sort start at once; the points follow on the copy stream in
    // chunks, each converted to niels entries as soon as it has landed; the bucket accumulation waits for the last chunk
    if ((rc = ctx->stage_in(scalars, n * 32))) return rc;
    if ((rc = ctx->reserve_out(32))) return rc;
    BBP_CUDA_OK(cudaEventRecord(ctx->ev_start, ctx->stream));            // scratch is free (previous work on the main stream done)
    BBP_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_start, 0));
    msm_shape sh = msm_engine::make_shape((uint32_t)n, (uint32_t)n, (uint32_t)n, false, 0, 0, 0);
    if ((rc = ctx->msm.run(sh, ctx->d_in, d_table, nullptr, ctx->d_out, 1))) return rc;   // scalar side, enqueued before the point copies
    BBP_CUDA_OK(cudaStreamWaitEvent(ctx->conv_stream, ctx->ev_start, 0));
    if (compressed) BBP_CUDA_OK(cudaMemsetAsync(d_valid, 1, 4, ctx->conv_stream));
    const size_t n_chunks = n >= (1u << 16) ? 8 : 1;
    const size_t chunk = ((n + n_chunks - 1) / n_chunks + BBP_NIELS_BATCH - 1) / BBP_NIELS_BATCH * BBP_NIELS_BATCH;
    size_t ci = 0;
    for (size_t off = 0; off < n; off += chunk, ci++) {
        size_t cnt = std::min(chunk, n - off);
        BBP_CUDA_OK(cudaMemcpyAsync(d_pts + off * pt_bytes, points + off * pt_bytes, cnt * pt_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        BBP_CUDA_OK(cudaEventRecord(ctx->ev_chunk[ci], ctx->copy_stream));
        BBP_CUDA_OK(cudaStreamWaitEvent(ctx->conv_stream, ctx->ev_chunk[ci], 0));
        if (compressed) {
            k_decompress_to_niels<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->conv_stream>>>((const uint32_t *)(d_pts + off * 32), d_table + off * 96,
                                                                                             (uint32_t)cnt, d_valid, nullptr);
        } else {
            size_t threads = (cnt + BBP_NIELS_BATCH - 1) / BBP_NIELS_BATCH;
            k_ext_to_niels<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->conv_stream>>>(d_pts + off * 128, d_table + off * 96, (uint32_t)cnt);
        }
        ctx->launches++;
    }
    BBP_CUDA_OK(cudaEventRecord(ctx->ev_table, ctx->conv_stream));
    ctx->msm.table_ready = ctx->ev_table;
    rc = ctx->msm.run(sh, ctx->d_in, d_table, nullptr, ctx->d_out, 2);
    ctx->msm.table_ready = nullptr;
    if (rc) { cudaStreamSynchronize(ctx->conv_stream); return rc; }
    if (compressed) BBP_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_table, 0));
    uint8_t res[32];
    BBP_CUDA_OK(cudaMemcpyAsync(res, ctx->d_out, 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (compressed) BBP_CUDA_OK(cudaMemcpyAsync(&h_valid, d_valid, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = wait_stream(ctx))) return rc;
    trace.mark("h2d+table+msm+d2h");
    if (!h_valid) return BBP_ERR_DECOMPRESS;   // optional_multiscalar_mul -> None; out untouched
    memcpy(out, res, 32);
    return BBP_OK;
}

int bbp_msm_vartime(bbp_ctx *ctx, const uint8_t *scalars, const uint8_t *points_ext, size_t n, uint8_t out[32]) {
    return msm_oneshot(ctx, scalars, points_ext, n, false, out);
}

int bbp_msm_optional(bbp_ctx *ctx, const uint8_t *scalars, const uint8_t *points_compressed, size_t n, uint8_t out[32]) {
    return msm_oneshot(ctx, scalars, points_compressed, n, true, out);
}

// ---------------------------------------------------------------- codecs
int bbp_decompress(bbp_ctx *ctx, const uint8_t *compressed, size_t n, uint8_t *out_ext, uint8_t *valid) {
    if (!ctx || !compressed || !out_ext || n == 0 || n > 0x7fffffffu) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    int rc = ctx->stage_in(compressed, n * 32);
    if (rc) return rc;
    size_t all_off = (n * 129 + 3) & ~(size_t)3;   // [n x 128 ext][n flags][pad][int all_valid]
    rc = ctx->reserve_out(all_off + 4);
    if (rc) return rc;
    uint8_t *d_flags = ctx->d_out + n * 128;
    int *d_all = (int *)(ctx->d_out + all_off);
    BBP_CUDA_OK(cudaMemsetAsync(d_all, 1, 4, ctx->stream));
    k_decompress_to_ext<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t *)ctx->d_in, ctx->d_out, (uint32_t)n, d_all, d_flags);
    ctx->launches++;
    BBP_CUDA_OK(cudaMemcpyAsync(out_ext, ctx->d_out, n * 128, cudaMemcpyDeviceToHost, ctx->stream));
    if (valid) BBP_CUDA_OK(cudaMemcpyAsync(valid, d_flags, n, cudaMemcpyDeviceToHost, ctx->stream));
    BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BBP_OK;
}

int bbp_compress(bbp_ctx *ctx, const uint8_t *points_ext, size_t n, uint8_t *out_compressed) {
    if (!ctx || !points_ext || !out_compressed || n == 0 || n > 0x7fffffffu) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    int rc = ctx->stage_in(points_ext, n * 128);
    if (rc) return rc;
    rc = ctx->reserve_out(n * 32);
    if (rc) return rc;
    k_compress<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_in, (uint32_t *)ctx->d_out, (uint32_t)n);
    ctx->launches++;
    BBP_CUDA_OK(cudaMemcpyAsync(out_compressed, ctx->d_out, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BBP_OK;
}

int bbp_from_uniform_bytes(bbp_ctx *ctx, const uint8_t *bytes64, size_t n, uint8_t *out_compressed) {
    if (!ctx || !bytes64 || !out_compressed || n == 0 || n > 0x7fffffffu) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    int rc = ctx->stage_in(bytes64, n * 64);
    if (rc) return rc;
    rc = ctx->reserve_out(n * 160);
    if (rc) return rc;
    k_from_uniform<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t *)ctx->d_in, ctx->d_out, (uint32_t)n);
    k_compress<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_out, (uint32_t *)(ctx->d_out + n * 128), (uint32_t)n);
    ctx->launches += 2;
    BBP_CUDA_OK(cudaMemcpyAsync(out_compressed, ctx->d_out + n * 128, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BBP_OK;
}

// ---------------------------------------------------------------- outer boundary: TLV codec + batched execution
struct bbp_wire_request {
    wire::request R;
    bool malformed = false;   // opcode 2 with a body the reference's readers reject: answered with 0x00 (futures/main.rs:95-101)
};

int bbp_wire_frame_len(const uint8_t *buf, size_t len, size_t *hdr_len, size_t *payload_len) {
    if (!buf || !hdr_len || !payload_len) return BBP_ERR_INPUT;
    size_t hdr;
    uint64_t pl;
    if (len == 0) return 0;
    if (!wire::tlv_header(buf, len, &hdr, &pl)) {
        int w = buf[0];
        return (w == 1 || w == 2 || w == 4 || w == 8) ? 0 : BBP_ERR_FORMAT;   // known tag, header not complete yet: need more bytes
    }
    if (pl > (1ull << 30)) return BBP_ERR_FORMAT;   // a request is a few kilobytes; refuse absurd lengths before buffering them
    *hdr_len = hdr;
    *payload_len = (size_t)pl;
    return len - hdr >= pl ? 1 : 0;
}

int bbp_wire_parse(const uint8_t *payload, size_t len, bbp_wire_request **out) {
    if (!payload || !out) return BBP_ERR_INPUT;
    *out = nullptr;
    bbp_wire_request *w = new (std::nothrow) bbp_wire_request();
    if (!w) return BBP_ERR_INPUT;
    int op = wire::parse_request(payload, len, w->R);
    if (op < 0 && len >= 1 && payload[0] == wire::OP_VERIFY) { w->malformed = true; op = wire::OP_VERIFY; }
    if (op <= 0) { delete w; return op < 0 ? BBP_ERR_FORMAT : 0; }
    *out = w;
    return op;
}
void bbp_wire_request_free(bbp_wire_request *r) { delete r; }
void bbp_wire_reply_free(uint8_t *reply) { free(reply); }

// n x len random bytes: SHAKE256(seed || LE64(index)) when a seed is given (tests), the OS generator otherwise
// randomness of request `index` of one bbp_wire_execute call: SHAKE256(seed || LE64(index)). seed is the caller's (tests:
// reproducible replies) or 32 bytes of OS entropy drawn ONCE per call (wire_os_seed) — opening the entropy device per
// request cost ~15 us each and was the server's throughput bound.
static void wire_entropy(const uint8_t *seed32, uint64_t index, uint8_t *out, size_t len) {
    keccak_sponge sp = shake256_new();
    sp.absorb(seed32, 32);
    uint8_t ix[8];
    for (int i = 0; i < 8; i++) ix[i] = (uint8_t)(index >> (8 * i));
    sp.absorb(ix, 8);
    sp.squeeze(out, len);
}
static bool wire_os_seed(uint8_t out[32]) {
    size_t got = 0;
    while (got < 32) {
        ssize_t r = getrandom(out + got, 32 - got, 0);
        if (r < 0) {
            if (errno == EINTR) continue;
            break;
        }
        got += (size_t)r;
    }
    if (got == 32) return true;
    FILE *f = fopen("/dev/urandom", "rb");   // kernels without the system call
    if (!f) return false;
    bool ok = fread(out, 1, 32, f) == 32;
    fclose(f);
    return ok;
}

static int wire_execute_impl(bbp_ctx *ctx, size_t n, bbp_wire_request *const *reqs, const uint8_t *seed32, uint8_t **replies, size_t *reply_lens);
int bbp_wire_execute(bbp_ctx *ctx, size_t n, bbp_wire_request *const *reqs, const uint8_t *seed32, uint8_t **replies, size_t *reply_lens) {
    if (!ctx || !reqs || !replies || !reply_lens || n == 0) return BBP_ERR_INPUT;
    for (size_t i = 0; i < n; i++) { replies[i] = nullptr; reply_lens[i] = 0; }
    int rc = wire_execute_impl(ctx, n, reqs, seed32, replies, reply_lens);
    if (rc)   // no partial results: whatever was encoded before the failure is released
        for (size_t i = 0; i < n; i++) { free(replies[i]); replies[i] = nullptr; reply_lens[i] = 0; }
    return rc;
}
static int wire_execute_impl(bbp_ctx *ctx, size_t n, bbp_wire_request *const *reqs, const uint8_t *seed32, uint8_t **replies, size_t *reply_lens) {
    cudaSetDevice(ctx->device);
    uint8_t os_seed[32];
    if (!seed32) {
        if (!wire_os_seed(os_seed)) return BBP_ERR_INPUT;
        seed32 = os_seed;
    }
    std::vector<prove_job> pj;
    std::vector<verify_job> vj;
    std::vector<size_t> pmap, vmap;
    for (size_t i = 0; i < n; i++) {
        replies[i] = nullptr;
        reply_lens[i] = 0;
        if (!reqs[i]) continue;
        const wire::request &R = reqs[i]->R;
        const size_t L = R.pub_list.size() / 32;
        if (R.opcode == wire::OP_PROVE) {
            prove_job J;
            // serde rejects non-canonical scalars (proof.rs:100-106); an empty list panics in the reference (gadgets.rs:103)
            if (L == 0 || !sc_from_canonical(J.d, R.scalars[0]) || !sc_from_canonical(J.k, R.scalars[1]) || !sc_from_canonical(J.y, R.scalars[2]) ||
                !sc_from_canonical(J.y_inv, R.scalars[3]) || !sc_from_canonical(J.q, R.scalars[4]) || !sc_from_canonical(J.z_img, R.scalars[5]) ||
                !sc_from_canonical(J.seed, R.scalars[6]))
                continue;
            J.pub_list.resize(L);
            for (size_t k = 0; k < L; k++) J.pub_list[k] = sc_from_bits(R.pub_list.data() + 32 * k);   // Bid::from (bid.rs:20-29)
            J.toggle = R.toggle;
            // the randomness the reference takes from thread_rng (proof.rs:53,64 and TranscriptRngBuilder::finalize)
            std::vector<uint8_t> rnd(64 * (4 + L) + 32);
            wire_entropy(seed32, i, rnd.data(), rnd.size());
            J.blindings.resize(4 + L);
            for (size_t k = 0; k < 4 + L; k++) J.blindings[k] = sc_from_wide(rnd.data() + 64 * k);   // Scalar::random
            memcpy(J.rng_seed, rnd.data() + 64 * (4 + L), 32);
            pmap.push_back(i);
            pj.push_back(std::move(J));
        } else if (R.opcode == wire::OP_VERIFY) {
            verify_job J;
            bool ok = !reqs[i]->malformed && sc_from_canonical(J.score, R.scalars[0]) && sc_from_canonical(J.z_img, R.scalars[1]) &&
                      sc_from_canonical(J.seed, R.scalars[2]);
            if (!ok) {
                wire::bytes rep = wire::verify_reply(false);
                replies[i] = (uint8_t *)malloc(rep.size());
                if (!replies[i]) return BBP_ERR_INPUT;
                memcpy(replies[i], rep.data(), rep.size());
                reply_lens[i] = rep.size();
                continue;
            }
            J.proof = R.proof.data(); J.proof_len = R.proof.size();
            J.commitments = R.commitments.data(); J.n_commitments = R.commitments.size() / 32;
            J.t_c = R.t_c.data(); J.n_t_c = R.t_c.size() / 32;
            J.pub_list.resize(L);
            for (size_t k = 0; k < L; k++) J.pub_list[k] = sc_from_bits(R.pub_list.data() + 32 * k);   // verify.rs:112-116
            wire_entropy(seed32, i, J.rng_seed, 32);
            vmap.push_back(i);
            vj.push_back(std::move(J));
        }
    }
    if (!pj.empty()) {
        int rc = prove_batch(ctx, pj);
        if (rc) return rc;
        for (size_t k = 0; k < pj.size(); k++) {
            const prove_job &J = pj[k];
            if (J.status) continue;   // prove-side error: nothing is written (futures/main.rs:15-25, 83-86)
            wire::bytes rep = wire::prove_reply(J.proof.data(), J.proof.size(), J.commitments.data(), J.commitments.size() / 32, J.t_c.data(), J.t_c.size() / 32);
            replies[pmap[k]] = (uint8_t *)malloc(rep.size());
            if (!replies[pmap[k]]) return BBP_ERR_INPUT;
            memcpy(replies[pmap[k]], rep.data(), rep.size());
            reply_lens[pmap[k]] = rep.size();
        }
    }
    if (!vj.empty()) {
        uint8_t batch_seed[32];
        wire_entropy(seed32, ~0ull, batch_seed, 32);
        int ok = 0;
        // one random linear combination for all pending verifications; on failure the library re-checks per request, so
        // every client gets the verdict Verify::verify would have given it
        int rc = verify_batch(ctx, vj, batch_seed, &ok, false, nullptr);
        if (rc) return rc;
        for (size_t k = 0; k < vj.size(); k++) {
            wire::bytes rep = wire::verify_reply(vj[k].status == 0);
            replies[vmap[k]] = (uint8_t *)malloc(rep.size());
            if (!replies[vmap[k]]) return BBP_ERR_INPUT;
            memcpy(replies[vmap[k]], rep.data(), rep.size());
            reply_lens[vmap[k]] = rep.size();
        }
    }
    return BBP_OK;
}

static int wire_copy_out(const wire::bytes &b, uint8_t *out, size_t *out_len) {
    const size_t cap = *out_len;
    *out_len = b.size();
    if (cap < b.size()) return BBP_ERR_INPUT;
    memcpy(out, b.data(), b.size());
    return BBP_OK;
}
int bbp_wire_encode_prove_request(const uint8_t *scalars7, const uint8_t *pub_list, size_t L, uint64_t toggle, uint8_t *out, size_t *out_len) {
    if (!scalars7 || (!pub_list && L) || !out || !out_len) return BBP_ERR_INPUT;
    uint8_t sc7[7][32];
    memcpy(sc7, scalars7, sizeof sc7);
    return wire_copy_out(wire::prove_request(sc7, pub_list, L, toggle), out, out_len);
}
int bbp_wire_encode_verify_request(const uint8_t *proof_blob, size_t blob_len, const uint8_t score[32], const uint8_t z_img[32], const uint8_t seed[32],
                                   const uint8_t *pub_list, size_t L, uint8_t *out, size_t *out_len) {
    if (!proof_blob || !score || !z_img || !seed || (!pub_list && L) || !out || !out_len) return BBP_ERR_INPUT;
    return wire_copy_out(wire::verify_request(proof_blob, blob_len, score, z_img, seed, pub_list, L), out, out_len);
}
int bbp_wire_encode_proof_blob(const uint8_t *proof, size_t proof_len, const uint8_t *commitments, size_t nc, const uint8_t *t_c, size_t nt, uint8_t *out,
                               size_t *out_len) {
    if (!proof || (!commitments && nc) || (!t_c && nt) || !out || !out_len) return BBP_ERR_INPUT;
    return wire_copy_out(wire::proof_blob(proof, proof_len, commitments, nc, t_c, nt), out, out_len);
}
int bbp_wire_decode_proof_blob(const uint8_t *blob, size_t blob_len, uint8_t *proof_out, size_t *proof_len, uint8_t *commitments_out, size_t *nc,
                               uint8_t *t_c_out, size_t *nt) {
    if (!blob || !proof_out || !proof_len || !commitments_out || !nc || !t_c_out || !nt) return BBP_ERR_INPUT;
    wire::bytes p, c, t;
    if (!wire::parse_proof_blob(blob, blob_len, p, c, t)) return BBP_ERR_FORMAT;
    if (p.size() > *proof_len || c.size() / 32 > *nc || t.size() / 32 > *nt) { *proof_len = p.size(); *nc = c.size() / 32; *nt = t.size() / 32; return BBP_ERR_INPUT; }
    memcpy(proof_out, p.data(), p.size()); memcpy(commitments_out, c.data(), c.size()); memcpy(t_c_out, t.data(), t.size());
    *proof_len = p.size(); *nc = c.size() / 32; *nt = t.size() / 32;
    return BBP_OK;
}

int bbp_set_ipp_shard(bbp_ctx *ctx, uint32_t rank, uint32_t world, bbp_allgather_fn allgather, void *user, int emulate) {
    if (!ctx || world == 0 || rank >= world || (world > 1 && !allgather) || emulate < 0 || emulate > 64 || world > 64) return BBP_ERR_INPUT;
    ctx->shard_rank = rank; ctx->shard_world = world; ctx->shard_allgather = allgather; ctx->shard_user = user;
    ctx->shard_emulate = world == 1 ? emulate : 0;
    return BBP_OK;
}

// ---------------------------------------------------------------- generic bulletproofs surface
struct bbp_transcript {
    merlin_transcript tr;
    bbp_transcript(const void *label, size_t n) : tr(label, n) {}
    explicit bbp_transcript(const merlin_transcript &o) : tr(o) {}
};
bbp_transcript *bbp_transcript_new(const uint8_t *label, size_t label_len) {
    if (!label && label_len) return nullptr;
    return new (std::nothrow) bbp_transcript(label ? (const void *)label : (const void *)"", label_len);
}
bbp_transcript *bbp_transcript_clone(const bbp_transcript *t) { return t ? new (std::nothrow) bbp_transcript(t->tr) : nullptr; }
void bbp_transcript_free(bbp_transcript *t) { delete t; }
int bbp_transcript_append_message(bbp_transcript *t, const uint8_t *label, size_t label_len, const uint8_t *msg, size_t msg_len) {
    if (!t || (!label && label_len) || (!msg && msg_len) || msg_len > 0xffffffffull) return BBP_ERR_INPUT;
    t->tr.append_message_l(label, label_len, msg, msg_len);
    return BBP_OK;
}
int bbp_transcript_append_u64(bbp_transcript *t, const uint8_t *label, size_t label_len, uint64_t x) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    return bbp_transcript_append_message(t, label, label_len, b, 8);
}
int bbp_transcript_challenge_bytes(bbp_transcript *t, const uint8_t *label, size_t label_len, uint8_t *out, size_t out_len) {
    if (!t || (!label && label_len) || (!out && out_len) || out_len > 0xffffffffull) return BBP_ERR_INPUT;
    t->tr.challenge_bytes_l(label, label_len, out, out_len);
    return BBP_OK;
}

static bool load_canonical(std::vector<sc> &out, const uint8_t *in, size_t n) {
    out.resize(n);
    for (size_t i = 0; i < n; i++)
        if (!sc_from_canonical(out[i], in + 32 * i)) return false;
    return true;
}

int bbp_cs_shape(const bbp_cs *cs, size_t out[5]) {
    if (!cs || !out || !cs->con_ptr || (cs->n_constraints && cs->con_ptr[cs->n_constraints] && (!cs->term_var || !cs->term_coeff))) return BBP_ERR_INPUT;
    generic_cs G;
    int rc = generic_cs_build(G, cs->n_multipliers, cs->n_commitments, cs->n_constraints, cs->con_ptr, cs->term_var, cs->term_coeff);
    if (rc) return rc;
    out[0] = G.tpl->n1; out[1] = G.tpl->q; out[2] = G.tpl->m; out[3] = next_pow2_u32(G.tpl->n1); out[4] = G.tpl->coef_table.size();
    return BBP_OK;
}

int bbp_r1cs_prove(bbp_ctx *ctx, bbp_transcript *t, const bbp_cs *cs, const uint8_t *a_L, const uint8_t *a_R, const uint8_t *a_O, const uint8_t *v,
                   const uint8_t *v_blinding, const uint8_t rng_seed[32], uint8_t *V_out, uint8_t *proof_out, size_t *proof_len) {
    if (!ctx || !t || !cs || !a_L || !a_R || !a_O || !rng_seed || !proof_out || !proof_len) return BBP_ERR_INPUT;
    if (cs->n_commitments && (!v || !v_blinding)) return BBP_ERR_INPUT;
    if (!cs->con_ptr || (cs->n_constraints && cs->con_ptr[cs->n_constraints] && (!cs->term_var || !cs->term_coeff))) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    generic_cs G;
    int rc = generic_cs_build(G, cs->n_multipliers, cs->n_commitments, cs->n_constraints, cs->con_ptr, cs->term_var, cs->term_coeff);
    if (rc) return rc;
    std::vector<sc> aL, aR, aO, vv, bl;
    if (!load_canonical(aL, a_L, cs->n_multipliers) || !load_canonical(aR, a_R, cs->n_multipliers) || !load_canonical(aO, a_O, cs->n_multipliers) ||
        !load_canonical(vv, v, cs->n_commitments) || !load_canonical(bl, v_blinding, cs->n_commitments))
        return BBP_ERR_FORMAT;
    std::vector<uint8_t> proof;
    int status = 0;
    rc = r1cs_prove_generic(ctx, t->tr, G, aL.data(), aR.data(), aO.data(), vv.data(), bl.data(), rng_seed, V_out, proof, &status);
    if (rc) return rc;
    if (status) return status;
    const size_t cap = *proof_len;
    *proof_len = proof.size();
    if (cap < proof.size()) return BBP_ERR_INPUT;
    memcpy(proof_out, proof.data(), proof.size());
    return BBP_OK;
}

int bbp_r1cs_verify(bbp_ctx *ctx, bbp_transcript *t, const bbp_cs *cs, const uint8_t *proof, size_t proof_len, const uint8_t *V,
                    const uint8_t rng_seed[32]) {
    if (!ctx || !t || !cs || (!proof && proof_len) || !rng_seed || (cs->n_commitments && !V)) return BBP_ERR_INPUT;
    if (!cs->con_ptr || (cs->n_constraints && cs->con_ptr[cs->n_constraints] && (!cs->term_var || !cs->term_coeff))) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    generic_cs G;
    int rc = generic_cs_build(G, cs->n_multipliers, cs->n_commitments, cs->n_constraints, cs->con_ptr, cs->term_var, cs->term_coeff);
    if (rc) return rc;
    int status = 0;
    rc = r1cs_verify_generic(ctx, t->tr, G, proof, proof_len, V, rng_seed, &status);
    return rc ? rc : status;
}

int bbp_ipp_create(bbp_ctx *ctx, bbp_transcript *t, const uint8_t w[32], const uint8_t *G_factors, const uint8_t *H_factors, const uint8_t *a,
                   const uint8_t *b, size_t n, uint8_t *proof_out, size_t *proof_len) {
    if (!ctx || !t || !w || !G_factors || !H_factors || !a || !b || !proof_out || !proof_len || n == 0 || n > (1u << 24)) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    sc ws;
    std::vector<sc> gf, hf, av, bv;
    if (!sc_from_canonical(ws, w) || !load_canonical(gf, G_factors, n) || !load_canonical(hf, H_factors, n) || !load_canonical(av, a, n) ||
        !load_canonical(bv, b, n))
        return BBP_ERR_FORMAT;
    std::vector<uint8_t> out;
    int rc = ipp_create_generic(ctx, t->tr, ws, gf.data(), hf.data(), av.data(), bv.data(), (uint32_t)n, out);
    if (rc) return rc;
    const size_t cap = *proof_len;
    *proof_len = out.size();
    if (cap < out.size()) return BBP_ERR_INPUT;
    memcpy(proof_out, out.data(), out.size());
    return BBP_OK;
}

// ---------------------------------------------------------------- test hooks
int bbp_test_fe(bbp_ctx *ctx, const uint8_t *a, const uint8_t *b, size_t n, int op, uint8_t *out) {
    if (!ctx || !a || !b || !out || n == 0) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    uint8_t *d = nullptr;
    BBP_CUDA_OK(cudaMalloc(&d, n * 96));
    cudaMemcpyAsync(d, a, n * 32, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d + n * 32, b, n * 32, cudaMemcpyHostToDevice, ctx->stream);
    k_test_fe<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t *)d, (const uint32_t *)(d + n * 32), (uint32_t *)(d + n * 64), (uint32_t)n, op);
    ctx->launches++;
    cudaMemcpyAsync(out, d + n * 64, n * 32, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    return e == cudaSuccess ? BBP_OK : BBP_ERR_CUDA;
}

int bbp_test_batch_weights(bbp_ctx *ctx, const uint8_t *r, size_t n, const uint8_t batch_seed[32], uint8_t *out) {
    if (!ctx || !r || !batch_seed || !out || n == 0 || n > (1u << 24)) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    std::vector<sc> chal(n * CH_N, sc_zero());
    for (size_t i = 0; i < n; i++)
        if (!sc_from_canonical(chal[i * CH_N + CH_R], r + 32 * i)) return BBP_ERR_FORMAT;
    uint8_t *d = nullptr;
    const size_t chal_bytes = n * CH_N * 32, dig_bytes = (n / 32 + 1) * 32;
    BBP_CUDA_OK(cudaMalloc(&d, chal_bytes + dig_bytes + 32 + n));
    uint8_t *d_dig = d + chal_bytes, *d_seed = d_dig + dig_bytes, *d_valid = d_seed + 32;
    cudaMemcpyAsync(d, chal.data(), chal_bytes, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_seed, batch_seed, 32, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(d_valid, 1, n, ctx->stream);
    k_batch_weight_digests<<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>((const sc *)d, (uint32_t)n, (uint64_t *)d_dig);
    k_batch_weights<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((sc *)d, d_valid, 1, (uint32_t)n, (const uint64_t *)d_dig, d_seed, 1u);
    ctx->launches += 2;
    cudaMemcpy2DAsync(out, 32, d + (size_t)CH_RHO * 32, (size_t)CH_N * 32, 32, n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    return e == cudaSuccess ? BBP_OK : BBP_ERR_CUDA;
}

int bbp_test_ge(bbp_ctx *ctx, const uint8_t *a_compressed, const uint8_t *b_compressed, size_t n, int op, uint8_t *out_compressed) {
    if (!ctx || !a_compressed || !b_compressed || !out_compressed || n == 0) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    uint8_t *d = nullptr;
    int *d_valid = nullptr;
    // layout: a_c | b_c | a_ext | b_ext | r_ext | r_c
    BBP_CUDA_OK(cudaMalloc(&d, n * (32 + 32 + 128 * 3 + 32)));
    BBP_CUDA_OK(cudaMalloc(&d_valid, 4));
    uint8_t *ac = d, *bc = d + n * 32, *ae = d + n * 64, *be = ae + n * 128, *re = be + n * 128, *rc_ = re + n * 128;
    int one = 1;
    cudaMemcpyAsync(d_valid, &one, 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(ac, a_compressed, n * 32, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(bc, b_compressed, n * 32, cudaMemcpyHostToDevice, ctx->stream);
    unsigned g = (unsigned)((n + 127) / 128);
    k_decompress_to_ext<<<g, 128, 0, ctx->stream>>>((const uint32_t *)ac, ae, (uint32_t)n, d_valid, nullptr);
    k_decompress_to_ext<<<g, 128, 0, ctx->stream>>>((const uint32_t *)bc, be, (uint32_t)n, d_valid, nullptr);
    k_test_ge<<<g, 128, 0, ctx->stream>>>(ae, be, re, (uint32_t)n, op);
    k_compress<<<g, 128, 0, ctx->stream>>>(re, (uint32_t *)rc_, (uint32_t)n);
    ctx->launches += 4;
    cudaMemcpyAsync(out_compressed, rc_, n * 32, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&one, d_valid, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d); cudaFree(d_valid);
    if (e != cudaSuccess) return BBP_ERR_CUDA;
    return one ? BBP_OK : BBP_ERR_DECOMPRESS;
}

// ---------------------------------------------------------------- blind-bid entry points
static sc load_scalar(const uint8_t *p) { return sc_from_bytes_mod_order(p); }

int bbp_set_proof_format(bbp_ctx *ctx, int versioned) {
    if (!ctx) return BBP_ERR_INPUT;
    proto_get(ctx)->proof_versioned = versioned ? 1 : 0;
    return BBP_OK;
}

int bbp_blindbid_prove_batch(bbp_ctx *ctx, size_t n, bbp_prove_req *reqs) {
    if (!ctx || !reqs || n == 0) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    std::vector<prove_job> jobs(n);
    for (size_t i = 0; i < n; i++) {
        bbp_prove_req &R = reqs[i];
        prove_job &J = jobs[i];
        R.status = BBP_ERR_INPUT;
        R.proof_len = 0;
        if (!R.d || !R.k || !R.y || !R.y_inv || !R.q || !R.z_img || !R.seed || !R.pub_list || !R.blindings || !R.rng_seed || !R.proof_out ||
            !R.commitments_out || !R.t_c_out || R.L == 0)
            continue;
        // any toggle is accepted, as in the reference (`i as u64 == toggle`, src/blindbid/proof.rs:60-66): an index beyond the
        // list makes every toggle bit zero and yields a proof that cannot verify, not an error.
        // d .. seed arrive through serde in the reference (src/blindbid/proof.rs:100-106), which rejects non-canonical scalars
        if (!sc_from_canonical(J.d, R.d) || !sc_from_canonical(J.k, R.k) || !sc_from_canonical(J.y, R.y) || !sc_from_canonical(J.y_inv, R.y_inv) ||
            !sc_from_canonical(J.q, R.q) || !sc_from_canonical(J.z_img, R.z_img) || !sc_from_canonical(J.seed, R.seed)) {
            R.status = BBP_ERR_FORMAT;
            continue;
        }
        J.pub_list.resize(R.L);
        for (size_t k = 0; k < R.L; k++) J.pub_list[k] = sc_from_bits(R.pub_list + 32 * k);   // Bid::from (src/blindbid/bid.rs:20-29)
        J.toggle = R.toggle;
        J.blindings.resize(4 + R.L);
        for (size_t k = 0; k < 4 + R.L; k++) J.blindings[k] = load_scalar(R.blindings + 32 * k);
        memcpy(J.rng_seed, R.rng_seed, 32);
        J.status = 1;   // marked runnable
    }
    std::vector<prove_job> run;
    std::vector<size_t> map;
    for (size_t i = 0; i < n; i++)
        if (jobs[i].status == 1) { map.push_back(i); run.push_back(std::move(jobs[i])); }
    if (!run.empty()) {
        int rc = prove_batch(ctx, run);
        if (rc) return rc;
    }
    for (size_t k = 0; k < run.size(); k++) {
        bbp_prove_req &R = reqs[map[k]];
        prove_job &J = run[k];
        R.status = J.status;
        if (J.status) continue;
        if (J.proof.size() > R.proof_cap) { R.status = BBP_ERR_INPUT; R.proof_len = J.proof.size(); continue; }
        memcpy(R.proof_out, J.proof.data(), J.proof.size());
        R.proof_len = J.proof.size();
        memcpy(R.commitments_out, J.commitments.data(), J.commitments.size());
        memcpy(R.t_c_out, J.t_c.data(), J.t_c.size());
    }
    return BBP_OK;
}

int bbp_blindbid_prove(bbp_ctx *ctx, const uint8_t d[32], const uint8_t k[32], const uint8_t y[32], const uint8_t y_inv[32], const uint8_t q[32],
                       const uint8_t z_img[32], const uint8_t seed[32], const uint8_t *pub_list, size_t L, uint64_t toggle, const uint8_t *blindings,
                       const uint8_t rng_seed[32], uint8_t *proof_out, size_t *proof_len, uint8_t *commitments_out, uint8_t *t_c_out) {
    if (!proof_len) return BBP_ERR_INPUT;
    bbp_prove_req R;
    memset(&R, 0, sizeof R);
    R.d = d; R.k = k; R.y = y; R.y_inv = y_inv; R.q = q; R.z_img = z_img; R.seed = seed; R.pub_list = pub_list; R.L = L; R.toggle = toggle;
    R.blindings = blindings; R.rng_seed = rng_seed; R.proof_out = proof_out; R.proof_cap = *proof_len; R.commitments_out = commitments_out; R.t_c_out = t_c_out;
    int rc = bbp_blindbid_prove_batch(ctx, 1, &R);
    if (rc) return rc;
    *proof_len = R.proof_len;
    return R.status;
}

static int load_verify_jobs(size_t n, bbp_verify_req *reqs, std::vector<verify_job> &jobs, std::vector<size_t> &map) {
    std::vector<verify_job> all(n);
    std::vector<uint8_t> good(n, 0);
    parallel_for(n, [&](size_t i) {
        bbp_verify_req &R = reqs[i];
        R.status = BBP_ERR_INPUT;
        if (!R.proof || !R.commitments || !R.t_c || !R.score || !R.z_img || !R.seed || (!R.pub_list && R.L) || !R.rng_seed) return;
        verify_job &J = all[i];
        J.proof = R.proof; J.proof_len = R.proof_len;                      // views: the caller's buffers outlive the call
        J.commitments = R.commitments; J.n_commitments = R.n_commitments;
        J.t_c = R.t_c; J.n_t_c = R.n_t_c;
        // serde-decoded in the reference (src/blindbid/verify.rs:100-102): canonical encodings only
        if (!sc_from_canonical(J.score, R.score) || !sc_from_canonical(J.z_img, R.z_img) || !sc_from_canonical(J.seed, R.seed)) { R.status = BBP_ERR_FORMAT; return; }
        J.pub_list.resize(R.L);
        for (size_t k = 0; k < R.L; k++) J.pub_list[k] = sc_from_bits(R.pub_list + 32 * k);   // src/blindbid/verify.rs:112-116
        memcpy(J.rng_seed, R.rng_seed, 32);
        good[i] = 1;
    });
    bool every = true;
    for (size_t i = 0; i < n; i++) every = every && good[i];
    if (every) {
        jobs = std::move(all);
        map.resize(n);
        for (size_t i = 0; i < n; i++) map[i] = i;
        return 0;
    }
    for (size_t i = 0; i < n; i++)
        if (good[i]) { map.push_back(i); jobs.push_back(std::move(all[i])); }
    return 0;
}

int bbp_blindbid_verify_each(bbp_ctx *ctx, size_t n, bbp_verify_req *reqs) {
    if (!ctx || !reqs || n == 0) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    std::vector<verify_job> jobs;
    std::vector<size_t> map;
    load_verify_jobs(n, reqs, jobs, map);
    if (!jobs.empty()) {
        int rc = verify_each(ctx, jobs);
        if (rc) return rc;
    }
    for (size_t k = 0; k < jobs.size(); k++) reqs[map[k]].status = jobs[k].status;
    return BBP_OK;
}

int bbp_blindbid_verify(bbp_ctx *ctx, const uint8_t *proof, size_t proof_len, const uint8_t *commitments, size_t n_commitments, const uint8_t *t_c,
                        size_t n_t_c, const uint8_t score[32], const uint8_t z_img[32], const uint8_t seed[32], const uint8_t *pub_list, size_t L,
                        const uint8_t rng_seed[32]) {
    bbp_verify_req R;
    memset(&R, 0, sizeof R);
    R.proof = proof; R.proof_len = proof_len; R.commitments = commitments; R.n_commitments = n_commitments; R.t_c = t_c; R.n_t_c = n_t_c;
    R.score = score; R.z_img = z_img; R.seed = seed; R.pub_list = pub_list; R.L = L; R.rng_seed = rng_seed;
    int rc = bbp_blindbid_verify_each(ctx, 1, &R);
    return rc ? rc : R.status;
}

int bbp_blindbid_verify_batch(bbp_ctx *ctx, size_t n, bbp_verify_req *reqs, const uint8_t batch_seed[32], int *all_ok) {
    if (!ctx || !reqs || n == 0 || !batch_seed) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    phase_trace tc("bbp_blindbid_verify_batch");
    std::vector<verify_job> jobs;
    std::vector<size_t> map;
    load_verify_jobs(n, reqs, jobs, map);
    tc.mark("load_requests");
    int ok = 1;
    if (!jobs.empty()) {
        int rc = verify_batch(ctx, jobs, batch_seed, &ok, false, nullptr);
        if (rc) return rc;
    }
    tc.mark("verify_batch");
    for (size_t k = 0; k < jobs.size(); k++) reqs[map[k]].status = jobs[k].status;
    if (jobs.size() != n) ok = 0;
    if (all_ok) *all_ok = ok;
    return BBP_OK;
}

int bbp_blindbid_verify_batch_partial(bbp_ctx *ctx, size_t n, bbp_verify_req *reqs, const uint8_t batch_seed[32], void *partial_ext_device, int *local_ok) {
    if (!ctx || !reqs || n == 0 || !batch_seed || !partial_ext_device) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    std::vector<verify_job> jobs;
    std::vector<size_t> map;
    load_verify_jobs(n, reqs, jobs, map);
    int ok = 1;
    // requests refused while loading (null pointers, non-canonical scalars) keep their status and clear the local flag;
    // the remaining ones still form this GPU's partial sum (the identity when none is left)
    int rc = verify_batch(ctx, jobs, batch_seed, &ok, true, (uint8_t *)partial_ext_device);
    if (rc) return rc;
    for (size_t k = 0; k < jobs.size(); k++) reqs[map[k]].status = jobs[k].status;
    if (jobs.size() != n) ok = 0;
    if (local_ok) *local_ok = ok;
    return BBP_OK;
}

int bbp_mimc_hash(const uint8_t left[32], const uint8_t right[32], uint8_t out[32]) {
    if (!left || !right || !out) return BBP_ERR_INPUT;
    sc_tobytes(out, mimc_hash(load_scalar(left), load_scalar(right)));
    return BBP_OK;
}

int bbp_mimc_constants(uint8_t out[90 * 32]) {
    if (!out) return BBP_ERR_INPUT;
    const std::vector<sc> &c = mimc_constants();
    for (size_t i = 0; i < c.size(); i++) sc_tobytes(out + 32 * i, c[i]);
    return BBP_OK;
}

int bbp_blindbid_circuit_shape(size_t n_commitments, size_t n_toggles, size_t out[3]) {
    if (!out || n_commitments < 4 || n_toggles < 1 || n_toggles > BLINDBID_MAX_TOGGLES || n_commitments > BLINDBID_MAX_COMMITMENTS) return BBP_ERR_INPUT;
    if (n_toggles <= 256) {   // small circuits are recorded for real (what the tests compare with the formulas and the oracle)
        std::shared_ptr<const circuit_template> t = blindbid_template((uint32_t)n_commitments, (uint32_t)n_toggles);
        out[0] = t->n1; out[1] = t->q; out[2] = t->m;
        return (t->n1 == blindbid_n1((uint32_t)n_toggles) && t->q == blindbid_q((uint32_t)n_toggles)) ? BBP_OK : BBP_ERR_INPUT;
    }
    out[0] = blindbid_n1((uint32_t)n_toggles); out[1] = blindbid_q((uint32_t)n_toggles); out[2] = n_commitments + n_toggles;
    return BBP_OK;
}

// test hook: Pedersen commitments v*B + r*B_blinding for n (value, blinding) pairs
int bbp_pedersen_commit(bbp_ctx *ctx, const uint8_t *values, const uint8_t *blindings, size_t n, uint8_t *out) {
    if (!ctx || !values || !blindings || !out || n == 0) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    int rc = proto_tables(ctx);
    if (rc) return rc;
    std::vector<sc> vals(2 * n);
    for (size_t i = 0; i < n; i++) { vals[2 * i] = load_scalar(values + 32 * i); vals[2 * i + 1] = load_scalar(blindings + 32 * i); }
    return pedersen_commit_host(ctx, vals.data(), n, out);
}

// n_slots MSMs over the resident generator table [B, B_blinding, G.., H..]; scalars = n_slots x slot_len x 32 B (host)
int bbp_msm_gens(bbp_ctx *ctx, const uint8_t *scalars, size_t slot_len, size_t n_slots, uint8_t *out) {
    if (!ctx || !scalars || !out || slot_len == 0 || n_slots == 0 || slot_len * n_slots > 0x7fffffffu) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    int rc = proto_tables(ctx);
    if (rc) return rc;
    if ((rc = ctx->stage_in(scalars, slot_len * n_slots * 32))) return rc;
    if ((rc = ctx->reserve_out(n_slots * 32))) return rc;
    if ((rc = msm_gens_device(ctx, (const sc *)ctx->d_in, (uint32_t)slot_len, (uint32_t)n_slots, ctx->d_out, nullptr))) return rc;
    BBP_CUDA_OK(cudaMemcpyAsync(out, ctx->d_out, n_slots * 32, cudaMemcpyDeviceToHost, ctx->stream));
    BBP_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BBP_OK;
}

// ---------------------------------------------------------------- aggregated range proofs (BASELINE config 5)
int bbp_rangeproof_prove_batch(bbp_ctx *ctx, size_t n_proofs, const uint64_t *values, const uint8_t *blindings, size_t m, size_t nbits,
                               const uint8_t *rng_seeds, uint8_t *proofs_out, size_t proof_stride, size_t *proof_len, uint8_t *commitments_out,
                               int *statuses) {
    if (!ctx || !values || !blindings || !rng_seeds || !proofs_out || !proof_len || !commitments_out || n_proofs == 0) return BBP_ERR_INPUT;
    if (!rp_params_ok(nbits, m)) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    std::vector<rp_prove_job> jobs(n_proofs);
    for (size_t i = 0; i < n_proofs; i++) {
        jobs[i].values.assign(values + i * m, values + (i + 1) * m);
        // a value that does not fit nbits is NOT rejected here (as upstream): the proof is produced and fails to verify
        jobs[i].blindings.resize(m);
        for (size_t j = 0; j < m; j++) jobs[i].blindings[j] = load_scalar(blindings + 32 * (i * m + j));
        memcpy(jobs[i].rng_seed, rng_seeds + 32 * i, 32);
    }
    int rc = rp_prove_group(ctx, jobs, (uint32_t)nbits);
    if (rc) return rc;
    size_t len = 0;
    for (size_t i = 0; i < n_proofs; i++) {
        if (statuses) statuses[i] = jobs[i].status;
        if (jobs[i].status) continue;
        len = jobs[i].proof.size();
        if (len > proof_stride) return BBP_ERR_INPUT;
        memcpy(proofs_out + i * proof_stride, jobs[i].proof.data(), len);
        memcpy(commitments_out + 32 * i * m, jobs[i].commitments.data(), 32 * m);
    }
    *proof_len = len;
    return BBP_OK;
}

int bbp_rangeproof_prove_multiple(bbp_ctx *ctx, const uint64_t *values, const uint8_t *blindings, size_t m, size_t nbits, const uint8_t rng_seed[32],
                                  uint8_t *proof_out, size_t *proof_len, uint8_t *commitments_out) {
    if (!proof_len) return BBP_ERR_INPUT;
    int st = 0;
    size_t cap = *proof_len;
    int rc = bbp_rangeproof_prove_batch(ctx, 1, values, blindings, m, nbits, rng_seed, proof_out, cap, proof_len, commitments_out, &st);
    return rc ? rc : st;
}

int bbp_rangeproof_verify_batch(bbp_ctx *ctx, size_t n_proofs, const uint8_t *proofs, size_t proof_stride, size_t proof_len, const uint8_t *commitments,
                                size_t m, size_t nbits, const uint8_t *rng_seeds, int *statuses) {
    if (!ctx || !proofs || !commitments || !rng_seeds || !statuses || n_proofs == 0 || proof_len > proof_stride) return BBP_ERR_INPUT;
    if (!rp_params_ok(nbits, m)) return BBP_ERR_INPUT;
    cudaSetDevice(ctx->device);
    std::vector<rp_verify_job> jobs(n_proofs);
    for (size_t i = 0; i < n_proofs; i++) {
        jobs[i].proof.assign(proofs + i * proof_stride, proofs + i * proof_stride + proof_len);
        jobs[i].commitments.assign(commitments + 32 * i * m, commitments + 32 * (i + 1) * m);
        memcpy(jobs[i].rng_seed, rng_seeds + 32 * i, 32);
    }
    int rc = rp_verify_group(ctx, jobs, (uint32_t)nbits, (uint32_t)m);
    if (rc) return rc;
    for (size_t i = 0; i < n_proofs; i++) statuses[i] = jobs[i].status;
    return BBP_OK;
}

int bbp_rangeproof_verify_multiple(bbp_ctx *ctx, const uint8_t *proof, size_t proof_len, const uint8_t *commitments, size_t m, size_t nbits,
                                   const uint8_t rng_seed[32]) {
    int st = 0;
    int rc = bbp_rangeproof_verify_batch(ctx, 1, proof, proof_len, proof_len, commitments, m, nbits, rng_seed, &st);
    return rc ? rc : st;
}

}  // extern "C"
