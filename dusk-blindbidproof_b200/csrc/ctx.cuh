// Backend context: one CUDA device + stream, resident generator tables, staging buffers, the MSM engine.
// The generator set is what generate_cs_transcript() builds on every request in the reference
// (src/blindbid/mod.rs:34-40: PedersenGens::default(), BulletproofGens::new(2048, 1)); here it is built once on the
// GPU (SHAKE256 stream on the host, Elligator maps + additions in k_from_uniform) and stays in HBM.
#pragma once
#include <vector>
#include "../../include/bbp.h"
#include "codec.cuh"
#include "keccak.h"
#include "msm.cuh"

namespace bbp { struct proto_state; void proto_release(proto_state *); }

struct bbp_points {
    bbp_ctx *ctx = nullptr;
    size_t n = 0;
    uint8_t *d_niels = nullptr;
};

struct bbp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // one-shot MSMs: point upload (copy_stream) and table build (conv_stream) overlap
    cudaStream_t conv_stream = nullptr;    // the scalar-side pipeline; uploads are never queued behind a conversion kernel
    cudaEvent_t ev_table = nullptr, ev_start = nullptr, ev_chunk[8] = {};
    cudaEvent_t ev_block = nullptr;   // cudaEventBlockingSync: waits that put the host thread to sleep (wait_stream)
    bbp::msm_engine msm;
    uint64_t launches = 0;
    // generators: index 0 = B, 1 = B_blinding, then G[party 0][0..cap), G[party 1][..), ..., then all H the same way, so
    // that the aggregated vectors bulletproofs iterates (party-major) are contiguous column ranges
    uint32_t gens_capacity = 0, party_capacity = 0;
    size_t n_gens = 0;
    uint8_t *d_gens_ext = nullptr;     // n_gens x 128 B
    uint8_t *d_gens_niels = nullptr;   // n_gens x 96 B
    uint8_t pc_compressed[64];
    // protocol layer state (templates, tables, scratch): protocol.cuh
    bbp::proto_state *proto = nullptr;
    // sibling contexts on the same device (own streams, scratch and tables) that large batched prove calls are split
    // over, one host thread each, so that one lane's host phases and round trips hide behind the others' kernels
    std::vector<bbp_ctx *> lanes;
    // sharded inner-product argument: this context owns the generator columns i = shard_rank (mod
    // shard_world); after every round's MSM the per-rank partial sums are exchanged through the caller's all-gather
    // (NCCL via torch.distributed in the Python host; stream-ordered on `stream`). shard_emulate: all shards computed here
    // one after the other, no collective (single-GPU tests of the partition).
    uint32_t shard_rank = 0, shard_world = 1;
    int shard_emulate = 0;
    bbp_allgather_fn shard_allgather = nullptr;
    void *shard_user = nullptr;
    // staging
    uint8_t *d_in = nullptr, *d_out = nullptr, *d_scratch = nullptr;
    size_t cap_in = 0, cap_out = 0, cap_scratch = 0;

    size_t gen_index(int which, uint32_t party, uint32_t i) const {
        return 2 + (which == 'H' ? (size_t)gens_capacity * party_capacity : 0) + (size_t)party * gens_capacity + i;
    }

    int stage_in(const void *host, size_t bytes) {
        if (bytes > cap_in) {
            cudaFree(d_in);
            d_in = nullptr; cap_in = 0;
            BBP_CUDA_OK(cudaMalloc(&d_in, bytes));
            cap_in = bytes;
        }
        BBP_CUDA_OK(cudaMemcpyAsync(d_in, host, bytes, cudaMemcpyHostToDevice, stream));
        return 0;
    }
    int reserve_scratch(size_t bytes) {
        if (bytes > cap_scratch) {
            cudaFree(d_scratch);
            d_scratch = nullptr; cap_scratch = 0;
            BBP_CUDA_OK(cudaMalloc(&d_scratch, bytes));
            cap_scratch = bytes;
        }
        return 0;
    }
    int reserve_out(size_t bytes) {
        if (bytes > cap_out) {
            cudaFree(d_out);
            d_out = nullptr; cap_out = 0;
            BBP_CUDA_OK(cudaMalloc(&d_out, bytes));
            cap_out = bytes;
        }
        return 0;
    }

    int init(uint32_t gens_cap, uint32_t party_cap) {
        using namespace bbp;
        BBP_CUDA_OK(cudaSetDevice(device));
        BBP_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        BBP_CUDA_OK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        BBP_CUDA_OK(cudaStreamCreateWithFlags(&conv_stream, cudaStreamNonBlocking));
        for (auto &e : ev_chunk) BBP_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        BBP_CUDA_OK(cudaEventCreateWithFlags(&ev_table, cudaEventDisableTiming));
        BBP_CUDA_OK(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
        BBP_CUDA_OK(cudaEventCreateWithFlags(&ev_block, cudaEventDisableTiming | cudaEventBlockingSync));
        msm.stream = stream;
        gens_capacity = gens_cap;
        party_capacity = gens_cap ? party_cap : 0;
        n_gens = 2 + (size_t)2 * gens_capacity * party_capacity;
        BBP_CUDA_OK(cudaMalloc(&d_gens_ext, n_gens * 128));
        BBP_CUDA_OK(cudaMalloc(&d_gens_niels, n_gens * 96));
        // uniform bytes for every hashed generator; slot 0 (B) is filled from the basepoint constant afterwards
        std::vector<uint8_t> uni(n_gens * 64, 0);
        // B_blinding = from_uniform_bytes(SHA3-512(compress(B))): needs compress(B) first
        uint8_t *d_tmp = nullptr;
        BBP_CUDA_OK(cudaMalloc(&d_tmp, 64));
        k_store_basepoint<<<1, 1, 0, stream>>>(d_gens_ext);
        k_compress<<<1, 128, 0, stream>>>(d_gens_ext, (uint32_t *)d_tmp, 1);
        BBP_CUDA_OK(cudaMemcpyAsync(pc_compressed, d_tmp, 32, cudaMemcpyDeviceToHost, stream));
        BBP_CUDA_OK(cudaStreamSynchronize(stream));
        sha3_512(uni.data() + 64, pc_compressed, 32);
        for (uint32_t j = 0; j < party_capacity; j++) {
            for (int hh = 0; hh < 2; hh++) {
                keccak_sponge sh = shake256_new();
                sh.absorb("GeneratorsChain", 15);
                uint8_t label[5] = {(uint8_t)(hh ? 'H' : 'G'), (uint8_t)j, (uint8_t)(j >> 8), (uint8_t)(j >> 16), (uint8_t)(j >> 24)};
                sh.absorb(label, 5);
                sh.squeeze(uni.data() + 64 * gen_index(hh ? 'H' : 'G', j, 0), (size_t)gens_capacity * 64);
            }
        }
        int rc = stage_in(uni.data(), uni.size());
        if (rc) return rc;
        // points 1 .. n_gens-1 from their uniform bytes
        uint32_t cnt = (uint32_t)(n_gens - 1);
        k_from_uniform<<<(cnt + 127) / 128, 128, 0, stream>>>((const uint32_t *)(d_in + 64), d_gens_ext + 128, cnt);
        uint32_t thr = (uint32_t)((n_gens + BBP_NIELS_BATCH - 1) / BBP_NIELS_BATCH);
        k_ext_to_niels<<<(thr + 127) / 128, 128, 0, stream>>>(d_gens_ext, d_gens_niels, (uint32_t)n_gens);
        k_compress<<<1, 128, 0, stream>>>(d_gens_ext + 128, (uint32_t *)d_tmp, 1);
        launches += 5;
        BBP_CUDA_OK(cudaMemcpyAsync(pc_compressed + 32, d_tmp, 32, cudaMemcpyDeviceToHost, stream));
        BBP_CUDA_OK(cudaStreamSynchronize(stream));
        cudaFree(d_tmp);
        return 0;
    }

    void destroy() {
        cudaSetDevice(device);
        for (bbp_ctx *l : lanes) { l->destroy(); delete l; }
        lanes.clear();
        if (stream) cudaStreamSynchronize(stream);
        msm.release();
        bbp::proto_release(proto);
        proto = nullptr;
        cudaFree(d_gens_ext); cudaFree(d_gens_niels); cudaFree(d_in); cudaFree(d_out); cudaFree(d_scratch);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (conv_stream) cudaStreamDestroy(conv_stream);
        for (auto &e : ev_chunk) if (e) { cudaEventDestroy(e); e = nullptr; }
        conv_stream = nullptr;
        if (ev_table) cudaEventDestroy(ev_table);
        if (ev_start) cudaEventDestroy(ev_start);
        if (ev_block) cudaEventDestroy(ev_block);
        ev_block = nullptr;
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr; copy_stream = nullptr; ev_table = ev_start = nullptr;
    }
};
