//! Safe host-side wrapper over the B200 backend. NOT compiled in the build image (no Rust toolchain).
//!
//! * [`Context::prove`] mirrors `Proof::prove` (reference `src/blindbid/proof.rs:36-91`),
//! * [`Context::verify`] / [`Context::verify_batch`] mirror `Verify::verify` (`src/blindbid/verify.rs:47-89`),
//! * [`Context::vartime_multiscalar_mul`] / [`Context::optional_multiscalar_mul`] mirror dalek's `VartimeMultiscalarMul`.
//!
//! The randomness the CPU path takes from `thread_rng` is drawn here and handed to the backend explicitly
//! (RNG contract, `include/bbp.h`), so proofs are exactly as random as before while staying reproducible in tests.
use bbp_sys as sys;
use curve25519_dalek::ristretto::CompressedRistretto;
use curve25519_dalek::scalar::Scalar;
use rand::RngCore;
use std::ptr;

pub mod r1cs;

#[derive(Debug)]
pub enum Error {
    InvalidGeneratorsLength,
    Format,
    Verification,
    Input,
    Decompress,
    Cuda,
    Other(i32),
}

pub(crate) fn check(rc: i32) -> Result<(), Error> {
    match rc {
        sys::BBP_OK => Ok(()),
        sys::BBP_ERR_INVALID_GENERATORS_LENGTH => Err(Error::InvalidGeneratorsLength),
        sys::BBP_ERR_FORMAT => Err(Error::Format),
        sys::BBP_ERR_VERIFICATION => Err(Error::Verification),
        sys::BBP_ERR_INPUT => Err(Error::Input),
        sys::BBP_ERR_DECOMPRESS => Err(Error::Decompress),
        sys::BBP_ERR_CUDA => Err(Error::Cuda),
        other => Err(Error::Other(other)),
    }
}

/// One GPU: generator tables resident, one stream. `Send` but not `Sync` — one context per worker thread.
pub struct Context {
    raw: *mut sys::bbp_ctx,
}
unsafe impl Send for Context {}

impl Context {
    pub(crate) fn raw(&self) -> *mut sys::bbp_ctx {
        self.raw
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::bbp_free(self.raw) }
    }
}

pub struct BlindBidProof {
    pub proof: Vec<u8>,
    pub commitments: Vec<CompressedRistretto>,
    pub t_c: Vec<CompressedRistretto>,
}

pub struct VerifyRequest<'a> {
    pub proof: &'a [u8],
    pub commitments: &'a [CompressedRistretto],
    pub t_c: &'a [CompressedRistretto],
    pub score: Scalar,
    pub z_img: Scalar,
    pub seed: Scalar,
    pub pub_list: &'a [Scalar],
}

fn flat_points(p: &[CompressedRistretto]) -> Vec<u8> {
    p.iter().flat_map(|c| c.as_bytes().to_vec()).collect()
}
fn flat_scalars(s: &[Scalar]) -> Vec<u8> {
    s.iter().flat_map(|x| x.as_bytes().to_vec()).collect()
}

impl Context {
    /// `generate_cs_transcript()` of the reference (`src/blindbid/mod.rs:34-40`), built once instead of per request.
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::bbp_init(&mut raw, device, 2048, 1) })?;
        Ok(Context { raw })
    }

    /// `Proof::prove(d, k, y, y_inv, q, z_img, seed, pub_list, toggle)`
    #[allow(clippy::too_many_arguments)]
    pub fn prove(&mut self, d: Scalar, k: Scalar, y: Scalar, y_inv: Scalar, q: Scalar, z_img: Scalar, seed: Scalar,
                 pub_list: &[Scalar], toggle: u64) -> Result<BlindBidProof, Error> {
        let l = pub_list.len();
        let mut rng = rand::thread_rng();
        let blindings = flat_scalars(&(0..4 + l).map(|_| Scalar::random(&mut rng)).collect::<Vec<_>>());
        let mut rng_seed = [0u8; 32];
        rng.fill_bytes(&mut rng_seed);
        let list = flat_scalars(pub_list);
        let (mut proof, mut comm, mut t_c) = (vec![0u8; 2048], vec![0u8; 4 * 32], vec![0u8; 32 * l]);
        let mut req = sys::bbp_prove_req {
            d: d.as_bytes().as_ptr(), k: k.as_bytes().as_ptr(), y: y.as_bytes().as_ptr(), y_inv: y_inv.as_bytes().as_ptr(),
            q: q.as_bytes().as_ptr(), z_img: z_img.as_bytes().as_ptr(), seed: seed.as_bytes().as_ptr(),
            pub_list: list.as_ptr(), l, toggle, blindings: blindings.as_ptr(), rng_seed: rng_seed.as_ptr(),
            proof_out: proof.as_mut_ptr(), proof_cap: proof.len(), proof_len: 0,
            commitments_out: comm.as_mut_ptr(), t_c_out: t_c.as_mut_ptr(), status: 0,
        };
        check(unsafe { sys::bbp_blindbid_prove_batch(self.raw, 1, &mut req) })?;
        check(req.status)?;
        proof.truncate(req.proof_len);
        Ok(BlindBidProof {
            proof,
            commitments: comm.chunks(32).map(CompressedRistretto::from_slice).collect(),
            t_c: t_c.chunks(32).map(CompressedRistretto::from_slice).collect(),
        })
    }

    fn verify_reqs(reqs: &[VerifyRequest], bufs: &mut Vec<(Vec<u8>, Vec<u8>, Vec<u8>, [u8; 32])>) -> Vec<sys::bbp_verify_req> {
        let mut rng = rand::thread_rng();
        for r in reqs {
            let mut s = [0u8; 32];
            rng.fill_bytes(&mut s);
            bufs.push((flat_points(r.commitments), flat_points(r.t_c), flat_scalars(r.pub_list), s));
        }
        reqs.iter().zip(bufs.iter()).map(|(r, b)| sys::bbp_verify_req {
            proof: r.proof.as_ptr(), proof_len: r.proof.len(),
            commitments: b.0.as_ptr(), n_commitments: r.commitments.len(),
            t_c: b.1.as_ptr(), n_t_c: r.t_c.len(),
            score: r.score.as_bytes().as_ptr(), z_img: r.z_img.as_bytes().as_ptr(), seed: r.seed.as_bytes().as_ptr(),
            pub_list: b.2.as_ptr(), l: r.pub_list.len(), rng_seed: b.3.as_ptr(), status: 0,
        }).collect()
    }

    /// `Verify::verify(&self)`
    pub fn verify(&mut self, req: &VerifyRequest) -> Result<(), Error> {
        let mut bufs = Vec::new();
        let mut raw = Self::verify_reqs(std::slice::from_ref(req), &mut bufs);
        check(unsafe { sys::bbp_blindbid_verify_each(self.raw, 1, raw.as_mut_ptr()) })?;
        check(raw[0].status)
    }

    /// Many requests, one combined mega-check; per-request results equal what `verify` returns for each.
    pub fn verify_batch(&mut self, reqs: &[VerifyRequest]) -> Result<Vec<Result<(), Error>>, Error> {
        let mut bufs = Vec::new();
        let mut raw = Self::verify_reqs(reqs, &mut bufs);
        let mut seed = [0u8; 32];
        rand::thread_rng().fill_bytes(&mut seed);
        let mut all_ok = 0;
        check(unsafe { sys::bbp_blindbid_verify_batch(self.raw, raw.len(), raw.as_mut_ptr(), seed.as_ptr(), &mut all_ok) })?;
        Ok(raw.iter().map(|r| check(r.status)).collect())
    }

    /// `VartimeMultiscalarMul::optional_multiscalar_mul` over compressed points: `None` if any point fails to decompress.
    pub fn optional_multiscalar_mul(&mut self, scalars: &[Scalar], points: &[CompressedRistretto]) -> Result<Option<CompressedRistretto>, Error> {
        assert_eq!(scalars.len(), points.len());
        let (s, p) = (flat_scalars(scalars), flat_points(points));
        let mut out = [0u8; 32];
        match unsafe { sys::bbp_msm_optional(self.raw, s.as_ptr(), p.as_ptr(), scalars.len(), out.as_mut_ptr()) } {
            sys::BBP_OK => Ok(Some(CompressedRistretto(out))),
            sys::BBP_ERR_DECOMPRESS => Ok(None),
            rc => check(rc).map(|_| None),
        }
    }

    /// `VartimeMultiscalarMul::vartime_multiscalar_mul`: points are valid encodings by construction here.
    pub fn vartime_multiscalar_mul(&mut self, scalars: &[Scalar], points: &[CompressedRistretto]) -> Result<CompressedRistretto, Error> {
        self.optional_multiscalar_mul(scalars, points)?.ok_or(Error::Decompress)
    }
}
