/* bbp.h — C ABI of the B200 blind-bid Bulletproofs backend (libbbp_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path. Every entry point names the reference
 * interface it stands in for; INTEGRATION.md shows the Rust FFI binding a maintainer would add for each.
 *
 * Conventions
 *   - scalars: 32 bytes little-endian. Unless stated otherwise any 256-bit value is accepted and reduced mod l,
 *     which is what dalek arithmetic does with `Scalar::from_bits` values (src/blindbid/bid.rs:27).
 *   - points: 32-byte compressed Ristretto (CompressedRistretto::to_bytes), or "extended": 128 bytes = X,Y,Z,T as four
 *     32-byte little-endian field elements (what an in-memory RistrettoPoint holds).
 *   - all buffers are caller owned HOST memory unless the name says `_device`; outputs are written only on success.
 *   - return value: 0 / BBP_OK on success, a negative bbp_status otherwise. Nothing aborts or throws across the ABI.
 *   - a context owns one CUDA device, one stream, the resident generator tables and all scratch memory. Calls on one
 *     context are serialised by the caller; distinct contexts are independent. There is no CPU fallback: without a
 *     CUDA device bbp_init fails with BBP_ERR_CUDA.
 */
#ifndef BBP_H
#define BBP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bbp_status {
    BBP_OK = 0,
    BBP_ERR_INVALID_GENERATORS_LENGTH = -1, /* R1CSError::InvalidGeneratorsLength */
    BBP_ERR_FORMAT = -2,                    /* R1CSError::FormatError */
    BBP_ERR_VERIFICATION = -3,              /* R1CSError::VerificationError */
    BBP_ERR_INPUT = -10,                    /* malformed arguments (the reference would panic: SURVEY.md §5) */
    BBP_ERR_DECOMPRESS = -11,               /* optional_multiscalar_mul -> None */
    BBP_ERR_CUDA = -100,
    BBP_ERR_NCCL = -101
} bbp_status;

typedef struct bbp_ctx bbp_ctx;
typedef struct bbp_points bbp_points; /* device-resident base table (affine niels, 96 B per point) */

/* ---- context ---------------------------------------------------------------------------------------------------- */
/* Creates a context on CUDA device `device`. Builds PedersenGens::default() and BulletproofGens::new(gens_capacity,
 * party_capacity) on the GPU once and keeps them resident (the reference rebuilds them on every request:
 * generate_cs_transcript(), src/blindbid/mod.rs:34-40). gens_capacity = 0 skips the generator build. */
int bbp_init(bbp_ctx **out, int device, uint32_t gens_capacity, uint32_t party_capacity);
void bbp_free(bbp_ctx *ctx);
/* number of kernels launched by this context so far (bench.py reports it as gpu_launches) */
uint64_t bbp_launch_count(const bbp_ctx *ctx);
/* the CUDA stream all work of this context is enqueued on (cudaStream_t as an integer), for event timing */
uint64_t bbp_stream(const bbp_ctx *ctx);
int bbp_sync(bbp_ctx *ctx);   /* waits for the context's stream and for those of its lanes */
/* Lanes: sibling contexts on the same device, created on first use, owned by `ctx` (freed with it; never pass one to
 * bbp_free). k = 0 is `ctx` itself, k <= 7. Each lane has its own stream, MSM engine and scratch, so independent calls
 * issued on different lanes overlap on the GPU: the latency-bound tail of one MSM (bucket reduction, Horner chain) runs
 * beside the bucket accumulation of the next, and one call's host<->device copies beside another's kernels. Read-only
 * inputs (bbp_points tables, device scalar buffers) may be shared between lanes once the call that produced them has
 * completed (bbp_sync). The batched prove / verify entry points use the same lanes internally for batches above 1024
 * requests. One host thread per lane at a time (no reference counterpart: the reference is one CPU thread per request,
 * src/futures/main.rs:52-62). */
int bbp_lane(bbp_ctx *ctx, uint32_t k, bbp_ctx **lane);

/* ---- measurement hooks (bench.py): per-stage MSM timing with CUDA events on the context stream, the MSM plan for a
 * size, and the measured integer-multiply ceiling of the device (no reference counterpart) */
int bbp_set_profiling(bbp_ctx *ctx, int on);
/* ms[0..7): recode, scans, scatter+task table, bucket accumulation, bucket reduction level 1, merge levels, combine+compress */
int bbp_msm_stage_ms(bbp_ctx *ctx, float *ms, size_t n_stages);
/* out = {window bits c, windows W, max entries per task S, buckets per window B} chosen for an n-point MSM */
int bbp_msm_plan(size_t n, uint32_t out[4]);
/* 32x32+64 multiply-accumulates (IMAD.WIDE.U32), all SMs, 8 independent chains per thread, full occupancy: sustained rate
 * per second (a pure multiplier loop runs power-capped at ~1.45 GHz) and the rate per SM clock (from clock64), which is the
 * architectural issue rate and does not depend on the clock the chip happened to hold */
int bbp_int_peak(bbp_ctx *ctx, double *wide_mads_per_s, double *wide_mads_per_clk_per_sm);
/* the same loop written as the mad.lo.cc.u32 / madc.hi.u32 pairs of the field multiplier (one pair = one 32x32+64 product) */
int bbp_int_peak_pairs(bbp_ctx *ctx, double *pairs_per_s, double *pairs_per_clk_per_sm);

/* ---- generators (bulletproofs PedersenGens / BulletproofGens, used at src/blindbid/mod.rs:35-36) ------------------- */
/* compressed B, B_blinding */
int bbp_pedersen_gens(bbp_ctx *ctx, uint8_t B[32], uint8_t B_blinding[32]);
/* compressed G or H generators (which = 'G' or 'H') of one party: count x 32 bytes starting at index first */
int bbp_bulletproof_gens(bbp_ctx *ctx, int which, uint32_t party, uint32_t first, uint32_t count, uint8_t *out);

/* ---- base tables -------------------------------------------------------------------------------------------------- */
/* Vec<CompressedRistretto> -> decompressed resident bases. *all_valid = 0 if any encoding is invalid (those entries
 * become the identity); mirrors the Option<RistrettoPoint> stream fed to optional_multiscalar_mul. */
int bbp_points_from_compressed(bbp_ctx *ctx, const uint8_t *points, size_t n, bbp_points **out, int *all_valid);
/* Vec<RistrettoPoint> (extended coordinates) -> resident bases */
int bbp_points_from_extended(bbp_ctx *ctx, const uint8_t *points_ext, size_t n, bbp_points **out);
size_t bbp_points_len(const bbp_points *p);
void bbp_points_free(bbp_points *p);

/* ---- L0: curve25519-dalek trait surface --------------------------------------------------------- */
/* VartimeMultiscalarMul::vartime_multiscalar_mul(scalars, points) with resident points; out = compressed result */
int bbp_msm_points(bbp_ctx *ctx, const uint8_t *scalars, size_t n, const bbp_points *points, uint8_t out[32]);
/* same, one shot from extended host points (decompression not needed) */
int bbp_msm_vartime(bbp_ctx *ctx, const uint8_t *scalars, const uint8_t *points_ext, size_t n, uint8_t out[32]);
/* VartimeMultiscalarMul::optional_multiscalar_mul over compressed points: BBP_ERR_DECOMPRESS if any fails (None) */
int bbp_msm_optional(bbp_ctx *ctx, const uint8_t *scalars, const uint8_t *points_compressed, size_t n, uint8_t out[32]);
/* device-pointer variant for callers that already hold scalars in HBM (n x 32 B) and want the result left in HBM:
 * out_device = 32 B compressed and/or out_ext_device = 128 B extended partial sum (either may be NULL); fully
 * asynchronous on the context stream */
int bbp_msm_points_device(bbp_ctx *ctx, const void *scalars_device, size_t n, const bbp_points *points, void *out_device,
                          void *out_ext_device);
/* tail of a sharded MSM: sum of n <= 1024 extended partial sums (n x 128 B in HBM, e.g. the all-gather
 * of every GPU's out_ext_device) -> 32 B compressed in HBM; asynchronous on the context stream */
int bbp_sum_compress_device(bbp_ctx *ctx, const void *points_ext_device, size_t n, void *out_device);
/* batched form: n_slots independent MSMs of n_per_slot scalars each over the SAME resident bases (slot-major scalars);
 * out = n_slots x 32 B compressed */
int bbp_msm_points_batched(bbp_ctx *ctx, const uint8_t *scalars, size_t n_per_slot, size_t n_slots, const bbp_points *points, uint8_t *out);

/* ---- point codecs (CompressedRistretto::decompress / RistrettoPoint::compress / from_uniform_bytes) ------------------ */
int bbp_decompress(bbp_ctx *ctx, const uint8_t *compressed, size_t n, uint8_t *out_ext, uint8_t *valid /* n bytes, may be NULL */);
int bbp_compress(bbp_ctx *ctx, const uint8_t *points_ext, size_t n, uint8_t *out_compressed);
int bbp_from_uniform_bytes(bbp_ctx *ctx, const uint8_t *bytes64, size_t n, uint8_t *out_compressed);

/* ---- L3: blind-bid entry points (mirror Proof::prove, src/blindbid/proof.rs:36-91, and Verify::verify,
 * src/blindbid/verify.rs:47-89; bulletproofs Prover / Verifier / InnerProductProof underneath) ----
 * RNG contract replacing thread_rng (proof.rs:53,64 and the two finalize(&mut thread_rng()) sites inside bulletproofs):
 * the caller supplies the 4+L commitment blindings and the 32 "external randomness" bytes keyed into the TranscriptRng.
 * With those fixed every output byte is a deterministic function of the inputs.
 * Scalars d .. seed (and score, z_img, seed of a verification): canonical encodings only (< l), as the reference's serde
 * decoding enforces (proof.rs:100-106, verify.rs:100-102): BBP_ERR_FORMAT otherwise. pub_list items: Scalar::from_bits
 * semantics (bid.rs:27, verify.rs:115). Blindings: any 32 bytes, reduced mod l. toggle: any value; an index beyond the list
 * yields (as in the reference) a proof that does not verify. */
typedef struct bbp_prove_req {
    const uint8_t *d, *k, *y, *y_inv, *q, *z_img, *seed; /* 32 B each */
    const uint8_t *pub_list;                             /* L x 32 B */
    size_t L;                                            /* >= 1 (the reference panics on an empty list) */
    uint64_t toggle;                                     /* index of the bidder's own item (>= L: all toggle bits zero) */
    const uint8_t *blindings;                            /* (4 + L) x 32 B */
    const uint8_t *rng_seed;                             /* 32 B */
    uint8_t *proof_out;                                  /* R1CSProof::to_bytes() */
    size_t proof_cap, proof_len;                         /* capacity in, length out (1121 B for every legal L) */
    uint8_t *commitments_out;                            /* 4 x 32 B: V_d, V_k, V_y, V_y_inv */
    uint8_t *t_c_out;                                    /* L x 32 B: toggle commitments */
    int status;                                          /* out: BBP_OK or a bbp_status */
} bbp_prove_req;
typedef struct bbp_verify_req {
    const uint8_t *proof; size_t proof_len;
    const uint8_t *commitments; size_t n_commitments;    /* >= 4 */
    const uint8_t *t_c; size_t n_t_c;                    /* >= 1 */
    const uint8_t *score, *z_img, *seed;                 /* 32 B each */
    const uint8_t *pub_list; size_t L;                   /* L >= n_t_c */
    const uint8_t *rng_seed;                             /* 32 B */
    int status;                                          /* out: BBP_OK = accept, BBP_ERR_FORMAT / _VERIFICATION / _INVALID_GENERATORS_LENGTH */
} bbp_verify_req;
/* R1CSProof byte layout (SURVEY.md §8c risk R1): 1 = develop-branch form with a leading phase byte (default), 0 = legacy */
int bbp_set_proof_format(bbp_ctx *ctx, int versioned);
int bbp_blindbid_prove(bbp_ctx *ctx, const uint8_t d[32], const uint8_t k[32], const uint8_t y[32], const uint8_t y_inv[32], const uint8_t q[32],
                       const uint8_t z_img[32], const uint8_t seed[32], const uint8_t *pub_list, size_t L, uint64_t toggle, const uint8_t *blindings,
                       const uint8_t rng_seed[32], uint8_t *proof_out, size_t *proof_len, uint8_t *commitments_out, uint8_t *t_c_out);
/* n requests in one pass over the GPU (requests are grouped by L internally); per-request status in reqs[i].status */
int bbp_blindbid_prove_batch(bbp_ctx *ctx, size_t n, bbp_prove_req *reqs);
/* returns BBP_OK when the proof verifies, a negative bbp_status otherwise (the reference maps every Err to the byte 0x00) */
int bbp_blindbid_verify(bbp_ctx *ctx, const uint8_t *proof, size_t proof_len, const uint8_t *commitments, size_t n_commitments, const uint8_t *t_c,
                        size_t n_t_c, const uint8_t score[32], const uint8_t z_img[32], const uint8_t seed[32], const uint8_t *pub_list, size_t L,
                        const uint8_t rng_seed[32]);
/* n independent verifications in one pass (each with its own mega-check): reqs[i].status */
int bbp_blindbid_verify_each(bbp_ctx *ctx, size_t n, bbp_verify_req *reqs);
/* batch verification: ONE combined mega-check over all requests (random weights from a Merlin transcript over the proofs
 * and batch_seed); falls back to per-request checks when the combination fails, so reqs[i].status always equals what
 * bbp_blindbid_verify_each reports. *all_ok = 1 iff every request verifies. */
int bbp_blindbid_verify_batch(bbp_ctx *ctx, size_t n, bbp_verify_req *reqs, const uint8_t batch_seed[32], int *all_ok);
/* proof-range sharding across GPUs: runs only the combined pass over this GPU's requests and leaves its
 * partial sums (2 x 128 B extended points: static-base part, dynamic part) in HBM at partial_ext_device for the all-gather;
 * Requests rejected before the combination (malformed proof, identity point, point that fails to decompress) do not
 * enter the partial sum: they clear *local_ok and set reqs[i].status. The whole batch verifies iff EVERY rank reports
 * *local_ok = 1 AND the sum of all GPUs' partials is the identity (bbp_sharded_verdict_device, or bbp_sum_compress_device ->
 * 32 zero bytes). */
int bbp_blindbid_verify_batch_partial(bbp_ctx *ctx, size_t n, bbp_verify_req *reqs, const uint8_t batch_seed[32], void *partial_ext_device, int *local_ok);
/* the verdict after the all-gather, in one launch on the context's stream (no synchronisation): rows_device = world rows of
 * row_stride (>= 257) bytes, row r = rank r's 2 x 128 B partial sums followed by its local flag byte; out_device[0] = 1 iff
 * every flag is set and the sum of all 2 x world points is the identity (extended-coordinate test, no compression). */
int bbp_sharded_verdict_device(bbp_ctx *ctx, const void *rows_device, size_t world, size_t row_stride, void *out_device);
/* native MiMC-x^7 hash the circuit constrains (src/gadgets.rs:37-68) and its constants (src/blindbid/mod.rs:7-24): host helpers */
int bbp_mimc_hash(const uint8_t left[32], const uint8_t right[32], uint8_t out[32]);
int bbp_mimc_constants(uint8_t out[90 * 32]);
/* out = {multipliers n1, constraints q, commitments m} of the circuit for the given commitment / toggle counts */
int bbp_blindbid_circuit_shape(size_t n_commitments, size_t n_toggles, size_t out[3]);

/* ---- aggregated range proofs: bulletproofs RangeProof::prove_multiple / verify_multiple (BASELINE.json configs[4]; the
 * reference itself has no call site). m values of nbits bits each (nbits in {8,16,32,64}, m a power of two);
 * the context must have been created with bbp_init(device, nbits, parties >= m). Proof = 32 * (9 + 2 lg(nbits m)) bytes.
 * RNG contract: upstream hands every party an rng; here party j draws from the SHAKE256 stream of
 * (rng_seed || v_blinding_j || LE64(value_j) || LE32(j)), 64 bytes per scalar in upstream's per-party order (a_blinding,
 * s_blinding, s_L, s_R, then t_1 / t_2 blindings). The stream is keyed with the party's witness, so a seed reused with other
 * values does not repeat nonces; prover seeds must still be SECRET and should be fresh per proof. The verifier's merging
 * scalar c and batch weight come from transcript.build_rng().finalize(rng_seed) after the whole proof has been absorbed:
 * verifier seeds must be SECRET (unpredictable to whoever supplies the proofs); the same seed may serve every request.
 * Transcript label "bbp-rangeproof". */
int bbp_rangeproof_prove_multiple(bbp_ctx *ctx, const uint64_t *values, const uint8_t *blindings, size_t m, size_t nbits, const uint8_t rng_seed[32],
                                  uint8_t *proof_out, size_t *proof_len, uint8_t *commitments_out /* m x 32 */);
/* BBP_OK = accept */
int bbp_rangeproof_verify_multiple(bbp_ctx *ctx, const uint8_t *proof, size_t proof_len, const uint8_t *commitments, size_t m, size_t nbits,
                                   const uint8_t rng_seed[32]);
/* n_proofs aggregated proofs in one pass (inputs concatenated proof-major; proofs_out rows of proof_stride bytes) */
int bbp_rangeproof_prove_batch(bbp_ctx *ctx, size_t n_proofs, const uint64_t *values, const uint8_t *blindings, size_t m, size_t nbits,
                               const uint8_t *rng_seeds, uint8_t *proofs_out, size_t proof_stride, size_t *proof_len, uint8_t *commitments_out,
                               int *statuses);
int bbp_rangeproof_verify_batch(bbp_ctx *ctx, size_t n_proofs, const uint8_t *proofs, size_t proof_stride, size_t proof_len, const uint8_t *commitments,
                                size_t m, size_t nbits, const uint8_t *rng_seeds, int *statuses);

/* ---- L1 building blocks over the resident generators --------------------------------------------------------------- */
/* PedersenGens::commit: v*B + r*B_blinding for n (value, blinding) pairs; out = n x 32 B compressed */
int bbp_pedersen_commit(bbp_ctx *ctx, const uint8_t *values, const uint8_t *blindings, size_t n, uint8_t *out);
/* n_slots MSMs over the generator table in the order [B, B_blinding, G[0..cap), H[0..cap)] (party 0 first); scalars =
 * n_slots x slot_len x 32 B, column i multiplies generator i; out = n_slots x 32 B compressed */
int bbp_msm_gens(bbp_ctx *ctx, const uint8_t *scalars, size_t slot_len, size_t n_slots, uint8_t *out);

/* ---- L1: the bulletproofs surface under src/gadgets.rs and src/blindbid (third layer) ---------------------
 * A Rust `ConstraintSystem` implementation over this ABI (rust/bbp/src/r1cs.rs) records what the gadgets do and hands the
 * flattened circuit over; the prover / verifier are the ones the blind-bid entry points above run on. */

/* merlin::Transcript — replaces `Transcript::new(b"BlindBidProofGadget")` at /root/reference/src/blindbid/mod.rs:37 and the
 * `&mut Transcript` that Prover::new / Verifier::new borrow (src/blindbid/proof.rs:50, verify.rs:51). Labels and messages
 * are (pointer, length) pairs. */
typedef struct bbp_transcript bbp_transcript;
bbp_transcript *bbp_transcript_new(const uint8_t *label, size_t label_len);
bbp_transcript *bbp_transcript_clone(const bbp_transcript *t);
void bbp_transcript_free(bbp_transcript *t);
int bbp_transcript_append_message(bbp_transcript *t, const uint8_t *label, size_t label_len, const uint8_t *msg, size_t msg_len);
int bbp_transcript_append_u64(bbp_transcript *t, const uint8_t *label, size_t label_len, uint64_t x);
int bbp_transcript_challenge_bytes(bbp_transcript *t, const uint8_t *label, size_t label_len, uint8_t *out, size_t out_len);

/* Variable kinds of bulletproofs::r1cs::Variable, packed as kind << 28 | index */
#define BBP_VAR_COMMITTED 0u /* Variable::Committed(i): the i-th Prover::commit / Verifier::commit */
#define BBP_VAR_MUL_LEFT 1u  /* Variable::MultiplierLeft(i) */
#define BBP_VAR_MUL_RIGHT 2u /* Variable::MultiplierRight(i) */
#define BBP_VAR_MUL_OUT 3u   /* Variable::MultiplierOutput(i) */
#define BBP_VAR_ONE 4u       /* Variable::One() */

/* The circuit as `ConstraintSystem::{multiply, constrain}` leave it (call sites /root/reference/src/gadgets.rs:30, 53-62,
 * 80-85, 119-139): `multiply(l, r)` allocates multiplier i and contributes the two constraints l - L_i = 0, r - R_i = 0;
 * `constrain(lc)` contributes lc = 0. Constraint j owns terms con_ptr[j] .. con_ptr[j + 1]; term t is
 * (term_var[t], term_coeff[32 t .. 32 t + 32) canonical scalar). Constraint order = call order (it fixes the z powers). */
typedef struct bbp_cs {
    uint32_t n_multipliers;  /* >= 1 */
    uint32_t n_commitments;  /* m */
    uint32_t n_constraints;
    const uint32_t *con_ptr; /* n_constraints + 1 entries, con_ptr[0] = 0 */
    const uint32_t *term_var;
    const uint8_t *term_coeff;
} bbp_cs;

/* Prover::prove — /root/reference/src/blindbid/proof.rs:50-88 (Prover::new, commit x m, gadgets, prove). `t` is the transcript
 * as Transcript::new left it and is advanced exactly as the Rust prover advances its borrow. a_L / a_R / a_O: the
 * multiplier assignments (n_multipliers x 32 B each), v / v_blinding: the committed values and their blindings (m x 32 B),
 * rng_seed: the 32 external bytes of the RNG contract. V_out (m x 32 B, may be NULL) receives the commitments;
 * proof_out / *proof_len: capacity in, length out (BBP_ERR_INPUT with the needed length if too small).
 * Returns BBP_OK or an R1CSError mirror (BBP_ERR_INVALID_GENERATORS_LENGTH when the padded circuit exceeds the context). */
int bbp_r1cs_prove(bbp_ctx *ctx, bbp_transcript *t, const bbp_cs *cs, const uint8_t *a_L, const uint8_t *a_R, const uint8_t *a_O, const uint8_t *v,
                   const uint8_t *v_blinding, const uint8_t rng_seed[32], uint8_t *V_out, uint8_t *proof_out, size_t *proof_len);
/* Verifier::verify — /root/reference/src/blindbid/verify.rs:51-88. V: the m commitments. BBP_OK = accept;
 * BBP_ERR_FORMAT / BBP_ERR_VERIFICATION / BBP_ERR_INVALID_GENERATORS_LENGTH as the reference reports them. */
int bbp_r1cs_verify(bbp_ctx *ctx, bbp_transcript *t, const bbp_cs *cs, const uint8_t *proof, size_t proof_len, const uint8_t *V,
                    const uint8_t rng_seed[32]);
/* InnerProductProof::create ([UP] bulletproofs inner_product_proof.rs; reached through Prover::prove, proof.rs:88) with
 * Q = w * B, over the first n resident generators G, H (n a power of two <= gens_capacity). G_factors, H_factors, a, b:
 * n x 32 B canonical scalars. Output: L_0 R_0 .. a b = 32 (2 lg n + 2) bytes. */
/* host-only: validates a flattened circuit the way bbp_r1cs_prove / _verify will and reports its shape: out = {multipliers,
 * constraints, commitments, padded length n = next power of two, entries of the coefficient table (1 = every variable
 * coefficient is +-1)}. BBP_ERR_FORMAT: index out of range, unknown variable kind, non-canonical coefficient. */
int bbp_cs_shape(const bbp_cs *cs, size_t out[5]);
int bbp_ipp_create(bbp_ctx *ctx, bbp_transcript *t, const uint8_t w[32], const uint8_t *G_factors, const uint8_t *H_factors, const uint8_t *a,
                   const uint8_t *b, size_t n, uint8_t *proof_out, size_t *proof_len);

/* ---- outer (process) boundary: TLV request / reply codec and batched execution ---------------------------------------------
 * The reference serves one TLV frame per Unix-socket connection, payload byte 0 = opcode (/root/reference/src/futures/main.rs:64-110):
 * opcode 1 = prove (body: /root/reference/src/blindbid/proof.rs:97-114, reply :118-143), opcode 2 = verify (body:
 * /root/reference/src/blindbid/verify.rs:91-128, reply: one byte). csrc/server/bbp_server.cpp is the socket shell over these
 * entry points; it coalesces concurrent requests into one bbp_wire_execute call. The byte-level framing belongs to the crate
 * dusk-tlv @ 5be856b, whose source is absent: csrc/wire.h implements the recollected form and marks it UNPINNED. */
typedef struct bbp_wire_request bbp_wire_request;
/* is buf[0..len) a complete frame? 1 = yes (*hdr_len + *payload_len bytes), 0 = more bytes needed, BBP_ERR_FORMAT = bad tag / absurd length */
int bbp_wire_frame_len(const uint8_t *buf, size_t len, size_t *hdr_len, size_t *payload_len);
/* parses the PAYLOAD of a request frame. Returns the opcode (1 prove, 2 verify; a verify body the reference's readers would
 * reject still yields a request, answered 0x00 as at futures/main.rs:95-101), 0 for an unknown opcode and BBP_ERR_FORMAT for a
 * malformed prove body (both: the reference writes nothing). *out must be released with bbp_wire_request_free. */
int bbp_wire_parse(const uint8_t *payload, size_t len, bbp_wire_request **out);
void bbp_wire_request_free(bbp_wire_request *r);
/* Executes n parsed requests at once: every prove request in ONE bbp_blindbid_prove_batch, every verify request in ONE
 * bbp_blindbid_verify_batch (per-request verdicts). replies[i] = the complete reply frame (release with bbp_wire_reply_free), or
 * NULL when the reference would write nothing (prove-side error). Request i's blindings and RNG seed are SHAKE256(seed || LE64(i)):
 * seed32 = NULL draws seed from the OS once per call (getrandom), as the reference takes its randomness from thread_rng;
 * non-NULL uses the caller's 32 bytes (reproducible tests). */
int bbp_wire_execute(bbp_ctx *ctx, size_t n, bbp_wire_request *const *reqs, const uint8_t *seed32, uint8_t **replies, size_t *reply_lens);
void bbp_wire_reply_free(uint8_t *reply);
/* client-side encoders: the frames a client (the Go node) sends, and the proof blob inside a verify request / prove reply.
 * out_len: capacity in, length out (BBP_ERR_INPUT with the needed length if too small). scalars7 = d,k,y,y_inv,q,z_img,seed. */
int bbp_wire_encode_prove_request(const uint8_t *scalars7, const uint8_t *pub_list, size_t L, uint64_t toggle, uint8_t *out, size_t *out_len);
int bbp_wire_encode_verify_request(const uint8_t *proof_blob, size_t blob_len, const uint8_t score[32], const uint8_t z_img[32], const uint8_t seed[32],
                                   const uint8_t *pub_list, size_t L, uint8_t *out, size_t *out_len);
int bbp_wire_encode_proof_blob(const uint8_t *proof, size_t proof_len, const uint8_t *commitments, size_t nc, const uint8_t *t_c, size_t nt, uint8_t *out,
                               size_t *out_len);
int bbp_wire_decode_proof_blob(const uint8_t *blob, size_t blob_len, uint8_t *proof_out, size_t *proof_len, uint8_t *commitments_out, size_t *nc,
                               uint8_t *t_c_out, size_t *nt);

/* ---- sharded inner-product argument (BASELINE config 5 at N GPUs) ---------------------------------------
 * After bbp_set_ipp_shard every InnerProductProof::create this context runs (range proofs, R1CS proofs, bbp_ipp_create) is
 * computed cooperatively by `world` contexts, one per GPU, that are all driven with the SAME inputs: context `rank` owns the
 * generator columns i = rank (mod world) (a strided partition, so that the folding pairs (i, i + n / 2) never straddle ranks)
 * and after each round's MSM the ranks exchange their 2 x 128-byte partial sums per proof through `allgather` (called on the
 * calling host thread; the copy must be ordered on bbp_stream(ctx); 0 = ok). Every rank then sums the partials, compresses
 * L_j / R_j and derives the same challenge, so all ranks return byte-identical proofs — identical to the single-GPU proof.
 * world = 1 switches it off. emulate > 1 (with world = 1) computes that many shards one after the other on this GPU without
 * any collective: the single-GPU test of the partition. */
typedef int (*bbp_allgather_fn)(void *user, const void *send_device, void *recv_device, size_t bytes_per_rank);
int bbp_set_ipp_shard(bbp_ctx *ctx, uint32_t rank, uint32_t world, bbp_allgather_fn allgather, void *user, int emulate);

/* ---- unit-test hooks (field / group primitives evaluated on the GPU; tests/ compares them with the oracle) ---------- */
/* op: 0 mul, 1 add, 2 sub, 3 invert(a), 4 square(a), 5 neg(a); inputs are raw 256-bit limbs, output canonical */
int bbp_test_fe(bbp_ctx *ctx, const uint8_t *a, const uint8_t *b, size_t n, int op, uint8_t *out);
/* op: 0 add, 1 double(a), 2 mixed add via niels(b), 3 mixed sub via niels(b), 4 niels(b) -> extended; compressed output */
int bbp_test_ge(bbp_ctx *ctx, const uint8_t *a_compressed, const uint8_t *b_compressed, size_t n, int op, uint8_t *out_compressed);
/* weights of a batch verification's random linear combination (csrc/rng_kernels.cuh: k_batch_weights), from the n
 * per-request digests r (n x 32 B canonical scalars) and the caller's batch seed; out = n x 32 B canonical scalars */
int bbp_test_batch_weights(bbp_ctx *ctx, const uint8_t *r, size_t n, const uint8_t batch_seed[32], uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif /* BBP_H */
