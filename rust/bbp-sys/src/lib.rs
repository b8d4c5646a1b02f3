//! Raw bindings to `include/bbp.h`. One declaration per C entry point; see the header for the reference interface each
//! one stands in for. NOT compiled in the build image (no Rust toolchain) — kept in lock-step with the header by hand.
#![allow(non_camel_case_types)]
use std::os::raw::{c_int, c_void};

pub const BBP_OK: c_int = 0;
pub const BBP_ERR_INVALID_GENERATORS_LENGTH: c_int = -1;
pub const BBP_ERR_FORMAT: c_int = -2;
pub const BBP_ERR_VERIFICATION: c_int = -3;
pub const BBP_ERR_INPUT: c_int = -10;
pub const BBP_ERR_DECOMPRESS: c_int = -11;
pub const BBP_ERR_CUDA: c_int = -100;
pub const BBP_ERR_NCCL: c_int = -101;

#[repr(C)]
pub struct bbp_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct bbp_points {
    _private: [u8; 0],
}

#[repr(C)]
pub struct bbp_wire_request {
    _private: [u8; 0],
}
#[repr(C)]
pub struct bbp_transcript {
    _private: [u8; 0],
}

pub const BBP_VAR_COMMITTED: u32 = 0;
pub const BBP_VAR_MUL_LEFT: u32 = 1;
pub const BBP_VAR_MUL_RIGHT: u32 = 2;
pub const BBP_VAR_MUL_OUT: u32 = 3;
pub const BBP_VAR_ONE: u32 = 4;

/// Flattened constraint system (`include/bbp.h`: `bbp_cs`).
#[repr(C)]
pub struct bbp_cs {
    pub n_multipliers: u32,
    pub n_commitments: u32,
    pub n_constraints: u32,
    pub con_ptr: *const u32,
    pub term_var: *const u32,
    pub term_coeff: *const u8,
}

#[repr(C)]
pub struct bbp_prove_req {
    pub d: *const u8,
    pub k: *const u8,
    pub y: *const u8,
    pub y_inv: *const u8,
    pub q: *const u8,
    pub z_img: *const u8,
    pub seed: *const u8,
    pub pub_list: *const u8,
    pub l: usize,
    pub toggle: u64,
    pub blindings: *const u8,
    pub rng_seed: *const u8,
    pub proof_out: *mut u8,
    pub proof_cap: usize,
    pub proof_len: usize,
    pub commitments_out: *mut u8,
    pub t_c_out: *mut u8,
    pub status: c_int,
}

#[repr(C)]
pub struct bbp_verify_req {
    pub proof: *const u8,
    pub proof_len: usize,
    pub commitments: *const u8,
    pub n_commitments: usize,
    pub t_c: *const u8,
    pub n_t_c: usize,
    pub score: *const u8,
    pub z_img: *const u8,
    pub seed: *const u8,
    pub pub_list: *const u8,
    pub l: usize,
    pub rng_seed: *const u8,
    pub status: c_int,
}

extern "C" {
    // context
    pub fn bbp_init(out: *mut *mut bbp_ctx, device: c_int, gens_capacity: u32, party_capacity: u32) -> c_int;
    pub fn bbp_free(ctx: *mut bbp_ctx);
    pub fn bbp_sync(ctx: *mut bbp_ctx) -> c_int;
    pub fn bbp_lane(ctx: *mut bbp_ctx, k: u32, lane: *mut *mut bbp_ctx) -> c_int;
    pub fn bbp_launch_count(ctx: *const bbp_ctx) -> u64;
    pub fn bbp_stream(ctx: *const bbp_ctx) -> u64;
    pub fn bbp_set_proof_format(ctx: *mut bbp_ctx, versioned: c_int) -> c_int;
    // generators
    pub fn bbp_pedersen_gens(ctx: *mut bbp_ctx, b: *mut u8, b_blinding: *mut u8) -> c_int;
    pub fn bbp_bulletproof_gens(ctx: *mut bbp_ctx, which: c_int, party: u32, first: u32, count: u32, out: *mut u8) -> c_int;
    // base tables + MSM (dalek trait surface)
    pub fn bbp_points_from_compressed(ctx: *mut bbp_ctx, points: *const u8, n: usize, out: *mut *mut bbp_points, all_valid: *mut c_int) -> c_int;
    pub fn bbp_points_from_extended(ctx: *mut bbp_ctx, points_ext: *const u8, n: usize, out: *mut *mut bbp_points) -> c_int;
    pub fn bbp_points_len(p: *const bbp_points) -> usize;
    pub fn bbp_points_free(p: *mut bbp_points);
    pub fn bbp_msm_points(ctx: *mut bbp_ctx, scalars: *const u8, n: usize, points: *const bbp_points, out: *mut u8) -> c_int;
    pub fn bbp_msm_points_batched(ctx: *mut bbp_ctx, scalars: *const u8, n_per_slot: usize, n_slots: usize, points: *const bbp_points, out: *mut u8) -> c_int;
    pub fn bbp_msm_points_device(ctx: *mut bbp_ctx, scalars_device: *const c_void, n: usize, points: *const bbp_points, out_device: *mut c_void, out_ext_device: *mut c_void) -> c_int;
    pub fn bbp_sum_compress_device(ctx: *mut bbp_ctx, points_ext_device: *const c_void, n: usize, out_device: *mut c_void) -> c_int;
    pub fn bbp_msm_vartime(ctx: *mut bbp_ctx, scalars: *const u8, points_ext: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn bbp_msm_optional(ctx: *mut bbp_ctx, scalars: *const u8, points_compressed: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn bbp_msm_gens(ctx: *mut bbp_ctx, scalars: *const u8, slot_len: usize, n_slots: usize, out: *mut u8) -> c_int;
    pub fn bbp_pedersen_commit(ctx: *mut bbp_ctx, values: *const u8, blindings: *const u8, n: usize, out: *mut u8) -> c_int;
    // codecs
    pub fn bbp_decompress(ctx: *mut bbp_ctx, compressed: *const u8, n: usize, out_ext: *mut u8, valid: *mut u8) -> c_int;
    pub fn bbp_compress(ctx: *mut bbp_ctx, points_ext: *const u8, n: usize, out_compressed: *mut u8) -> c_int;
    pub fn bbp_from_uniform_bytes(ctx: *mut bbp_ctx, bytes64: *const u8, n: usize, out_compressed: *mut u8) -> c_int;
    // blind bid
    pub fn bbp_blindbid_prove_batch(ctx: *mut bbp_ctx, n: usize, reqs: *mut bbp_prove_req) -> c_int;
    pub fn bbp_blindbid_verify_each(ctx: *mut bbp_ctx, n: usize, reqs: *mut bbp_verify_req) -> c_int;
    pub fn bbp_blindbid_verify_batch(ctx: *mut bbp_ctx, n: usize, reqs: *mut bbp_verify_req, batch_seed: *const u8, all_ok: *mut c_int) -> c_int;
    pub fn bbp_blindbid_verify_batch_partial(ctx: *mut bbp_ctx, n: usize, reqs: *mut bbp_verify_req, batch_seed: *const u8, partial_ext_device: *mut c_void, local_ok: *mut c_int) -> c_int;
    pub fn bbp_sharded_verdict_device(ctx: *mut bbp_ctx, rows_device: *const c_void, world: usize, row_stride: usize, out_device: *mut c_void) -> c_int;
    pub fn bbp_mimc_hash(left: *const u8, right: *const u8, out: *mut u8) -> c_int;
    pub fn bbp_mimc_constants(out: *mut u8) -> c_int;
    pub fn bbp_blindbid_circuit_shape(n_commitments: usize, n_toggles: usize, out: *mut usize) -> c_int;
    // generic bulletproofs surface: transcript, R1CS prover / verifier, inner-product argument
    pub fn bbp_transcript_new(label: *const u8, label_len: usize) -> *mut bbp_transcript;
    pub fn bbp_transcript_clone(t: *const bbp_transcript) -> *mut bbp_transcript;
    pub fn bbp_transcript_free(t: *mut bbp_transcript);
    pub fn bbp_transcript_append_message(t: *mut bbp_transcript, label: *const u8, label_len: usize, msg: *const u8, msg_len: usize) -> c_int;
    pub fn bbp_transcript_append_u64(t: *mut bbp_transcript, label: *const u8, label_len: usize, x: u64) -> c_int;
    pub fn bbp_transcript_challenge_bytes(t: *mut bbp_transcript, label: *const u8, label_len: usize, out: *mut u8, out_len: usize) -> c_int;
    pub fn bbp_cs_shape(cs: *const bbp_cs, out: *mut usize) -> c_int;
    pub fn bbp_r1cs_prove(ctx: *mut bbp_ctx, t: *mut bbp_transcript, cs: *const bbp_cs, a_l: *const u8, a_r: *const u8, a_o: *const u8, v: *const u8, v_blinding: *const u8, rng_seed: *const u8, v_out: *mut u8, proof_out: *mut u8, proof_len: *mut usize) -> c_int;
    pub fn bbp_r1cs_verify(ctx: *mut bbp_ctx, t: *mut bbp_transcript, cs: *const bbp_cs, proof: *const u8, proof_len: usize, v: *const u8, rng_seed: *const u8) -> c_int;
    pub fn bbp_ipp_create(ctx: *mut bbp_ctx, t: *mut bbp_transcript, w: *const u8, g_factors: *const u8, h_factors: *const u8, a: *const u8, b: *const u8, n: usize, proof_out: *mut u8, proof_len: *mut usize) -> c_int;
    // outer boundary: TLV codec + batched execution
    pub fn bbp_wire_frame_len(buf: *const u8, len: usize, hdr_len: *mut usize, payload_len: *mut usize) -> c_int;
    pub fn bbp_wire_parse(payload: *const u8, len: usize, out: *mut *mut bbp_wire_request) -> c_int;
    pub fn bbp_wire_request_free(r: *mut bbp_wire_request);
    pub fn bbp_wire_execute(ctx: *mut bbp_ctx, n: usize, reqs: *const *mut bbp_wire_request, seed32: *const u8, replies: *mut *mut u8, reply_lens: *mut usize) -> c_int;
    pub fn bbp_wire_reply_free(reply: *mut u8);
    // aggregated range proofs
    pub fn bbp_rangeproof_prove_multiple(ctx: *mut bbp_ctx, values: *const u64, blindings: *const u8, m: usize, nbits: usize, rng_seed: *const u8, proof_out: *mut u8, proof_len: *mut usize, commitments_out: *mut u8) -> c_int;
    pub fn bbp_rangeproof_verify_multiple(ctx: *mut bbp_ctx, proof: *const u8, proof_len: usize, commitments: *const u8, m: usize, nbits: usize, rng_seed: *const u8) -> c_int;
}
