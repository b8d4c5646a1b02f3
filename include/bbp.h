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
int bbp_sync(bbp_ctx *ctx);

/* ---- measurement hooks (bench.py): per-stage MSM timing with CUDA events on the context stream, the MSM plan for a
 * size, and the measured integer-multiply ceiling of the device (no reference counterpart) */
int bbp_set_profiling(bbp_ctx *ctx, int on);
/* ms[0..7): recode, scans, scatter+task table, bucket accumulation, chunk reduce, window reduce, combine+compress */
int bbp_msm_stage_ms(bbp_ctx *ctx, float *ms, size_t n_stages);
/* out = {window bits c, windows W, max entries per task S, buckets per chunk CH} chosen for an n-point MSM */
int bbp_msm_plan(size_t n, uint32_t out[4]);
/* sustained 32x32->64 multiply-accumulates per second (IMAD.WIDE.U32), all SMs, 8 independent chains per thread */
int bbp_int_peak(bbp_ctx *ctx, double *wide_mads_per_s);

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

/* ---- unit-test hooks (field / group primitives evaluated on the GPU; tests/ compares them with the oracle) ---------- */
/* op: 0 mul, 1 add, 2 sub, 3 invert(a), 4 square(a), 5 neg(a); inputs are raw 256-bit limbs, output canonical */
int bbp_test_fe(bbp_ctx *ctx, const uint8_t *a, const uint8_t *b, size_t n, int op, uint8_t *out);
/* op: 0 add, 1 double(a), 2 mixed add via niels(b), 3 mixed sub via niels(b), 4 niels(b) -> extended; compressed output */
int bbp_test_ge(bbp_ctx *ctx, const uint8_t *a_compressed, const uint8_t *b_compressed, size_t n, int op, uint8_t *out_compressed);

#ifdef __cplusplus
}
#endif
#endif /* BBP_H */
