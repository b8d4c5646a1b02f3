//! `bulletproofs::r1cs`-shaped constraint system over the B200 backend. NOT compiled in the build image.
//!
//! The gadgets of the reference (`src/gadgets.rs`) are written against `ConstraintSystem::{multiply, constrain}` and
//! `LinearCombination`; this module provides the same trait surface. `Prover` / `Verifier` record what the gadgets do —
//! the multiplier assignments (prover) and the constraints in call order — and `prove` / `verify` hand the flattened
//! circuit to `bbp_r1cs_prove` / `bbp_r1cs_verify`, which run the same device prover / verifier as the blind-bid entry
//! points. With this module `src/gadgets.rs` compiles unchanged against `bbp::r1cs` instead of `bulletproofs::r1cs`:
//!
//! ```ignore
//! // src/blindbid/proof.rs:47-90, body unchanged except for the two `use` lines
//! let mut transcript = bbp::r1cs::Transcript::new(b"BlindBidProofGadget");
//! let mut prover = bbp::r1cs::Prover::new(&ctx, &mut transcript);
//! let (commit, var) = prover.commit(d, Scalar::random(&mut rng));
//! gadgets::proof_gadget(&mut prover, ...);
//! let proof = prover.prove(&mut rng)?;
//! ```
use crate::{check, Context, Error};
use bbp_sys as sys;
use curve25519_dalek::ristretto::CompressedRistretto;
use curve25519_dalek::scalar::Scalar;
use rand::RngCore;
use std::ops::{Add, Neg, Sub};

/// merlin::Transcript over `bbp_transcript_*`.
pub struct Transcript {
    raw: *mut sys::bbp_transcript,
}
impl Transcript {
    pub fn new(label: &'static [u8]) -> Self {
        Transcript { raw: unsafe { sys::bbp_transcript_new(label.as_ptr(), label.len()) } }
    }
    pub fn append_message(&mut self, label: &'static [u8], msg: &[u8]) {
        unsafe { sys::bbp_transcript_append_message(self.raw, label.as_ptr(), label.len(), msg.as_ptr(), msg.len()) };
    }
    pub fn append_u64(&mut self, label: &'static [u8], x: u64) {
        unsafe { sys::bbp_transcript_append_u64(self.raw, label.as_ptr(), label.len(), x) };
    }
    pub fn challenge_bytes(&mut self, label: &'static [u8], dest: &mut [u8]) {
        unsafe { sys::bbp_transcript_challenge_bytes(self.raw, label.as_ptr(), label.len(), dest.as_mut_ptr(), dest.len()) };
    }
}
impl Clone for Transcript {
    fn clone(&self) -> Self {
        Transcript { raw: unsafe { sys::bbp_transcript_clone(self.raw) } }
    }
}
impl Drop for Transcript {
    fn drop(&mut self) {
        unsafe { sys::bbp_transcript_free(self.raw) }
    }
}

/// bulletproofs::r1cs::Variable
#[derive(Copy, Clone, Debug, PartialEq)]
pub enum Variable {
    Committed(usize),
    MultiplierLeft(usize),
    MultiplierRight(usize),
    MultiplierOutput(usize),
    One(),
}
impl Variable {
    fn pack(self) -> u32 {
        let (kind, idx) = match self {
            Variable::Committed(i) => (sys::BBP_VAR_COMMITTED, i),
            Variable::MultiplierLeft(i) => (sys::BBP_VAR_MUL_LEFT, i),
            Variable::MultiplierRight(i) => (sys::BBP_VAR_MUL_RIGHT, i),
            Variable::MultiplierOutput(i) => (sys::BBP_VAR_MUL_OUT, i),
            Variable::One() => (sys::BBP_VAR_ONE, 0),
        };
        (kind << 28) | (idx as u32)
    }
}

/// bulletproofs::r1cs::LinearCombination: a term list; `+` / `-` concatenate (no deduplication), as upstream.
#[derive(Clone, Debug, Default)]
pub struct LinearCombination {
    pub terms: Vec<(Variable, Scalar)>,
}
impl From<Variable> for LinearCombination {
    fn from(v: Variable) -> Self {
        LinearCombination { terms: vec![(v, Scalar::one())] }
    }
}
impl From<Scalar> for LinearCombination {
    fn from(s: Scalar) -> Self {
        LinearCombination { terms: vec![(Variable::One(), s)] }
    }
}
impl<T: Into<LinearCombination>> Add<T> for LinearCombination {
    type Output = Self;
    fn add(mut self, rhs: T) -> Self {
        self.terms.extend(rhs.into().terms);
        self
    }
}
impl<T: Into<LinearCombination>> Sub<T> for LinearCombination {
    type Output = Self;
    fn sub(mut self, rhs: T) -> Self {
        self.terms.extend(rhs.into().terms.into_iter().map(|(v, s)| (v, -s)));
        self
    }
}
impl Neg for LinearCombination {
    type Output = Self;
    fn neg(mut self) -> Self {
        for (_, s) in self.terms.iter_mut() {
            *s = -*s;
        }
        self
    }
}

/// bulletproofs::r1cs::ConstraintSystem (the two methods the reference's gadgets call: `src/gadgets.rs:30,53`).
pub trait ConstraintSystem {
    fn multiply(&mut self, left: LinearCombination, right: LinearCombination) -> (Variable, Variable, Variable);
    fn constrain(&mut self, lc: LinearCombination);
}

#[derive(Default)]
struct Recorded {
    n_multipliers: usize,
    con_ptr: Vec<u32>,
    term_var: Vec<u32>,
    term_coeff: Vec<u8>,
}
impl Recorded {
    fn new() -> Self {
        Recorded { n_multipliers: 0, con_ptr: vec![0], term_var: Vec::new(), term_coeff: Vec::new() }
    }
    fn push(&mut self, lc: &LinearCombination) {
        for (v, s) in &lc.terms {
            self.term_var.push(v.pack());
            self.term_coeff.extend_from_slice(s.as_bytes());
        }
        self.con_ptr.push(self.term_var.len() as u32);
    }
    fn as_cs(&self, n_commitments: usize) -> sys::bbp_cs {
        sys::bbp_cs {
            n_multipliers: self.n_multipliers as u32,
            n_commitments: n_commitments as u32,
            n_constraints: (self.con_ptr.len() - 1) as u32,
            con_ptr: self.con_ptr.as_ptr(),
            term_var: self.term_var.as_ptr(),
            term_coeff: self.term_coeff.as_ptr(),
        }
    }
}

/// bulletproofs::r1cs::Prover — `src/blindbid/proof.rs:50-88`.
pub struct Prover<'a> {
    ctx: &'a Context,
    transcript: &'a mut Transcript,
    rec: Recorded,
    a_l: Vec<Scalar>,
    a_r: Vec<Scalar>,
    a_o: Vec<Scalar>,
    v: Vec<Scalar>,
    v_blinding: Vec<Scalar>,
}
impl<'a> Prover<'a> {
    pub fn new(ctx: &'a Context, transcript: &'a mut Transcript) -> Self {
        Prover { ctx, transcript, rec: Recorded::new(), a_l: vec![], a_r: vec![], a_o: vec![], v: vec![], v_blinding: vec![] }
    }
    /// Prover::commit (`proof.rs:57`): the commitment is returned at once (one fixed-base launch); the transcript absorbs
    /// it inside `prove`, in the same position as upstream (commitments never interleave with other transcript writes).
    pub fn commit(&mut self, v: Scalar, v_blinding: Scalar) -> Result<(CompressedRistretto, Variable), Error> {
        let mut out = [0u8; 32];
        check(unsafe { sys::bbp_pedersen_commit(self.ctx.raw(), v.as_bytes().as_ptr(), v_blinding.as_bytes().as_ptr(), 1, out.as_mut_ptr()) })?;
        let i = self.v.len();
        self.v.push(v);
        self.v_blinding.push(v_blinding);
        Ok((CompressedRistretto(out), Variable::Committed(i)))
    }
    fn eval(&self, lc: &LinearCombination) -> Scalar {
        lc.terms.iter().fold(Scalar::zero(), |acc, (var, c)| {
            acc + c * match var {
                Variable::Committed(i) => self.v[*i],
                Variable::MultiplierLeft(i) => self.a_l[*i],
                Variable::MultiplierRight(i) => self.a_r[*i],
                Variable::MultiplierOutput(i) => self.a_o[*i],
                Variable::One() => Scalar::one(),
            }
        })
    }
    /// Prover::prove (`proof.rs:88`): R1CSProof::to_bytes of the result. The 32 external RNG bytes upstream takes from
    /// `thread_rng` inside `TranscriptRngBuilder::finalize` are drawn here.
    pub fn prove<R: RngCore>(self, rng: &mut R) -> Result<Vec<u8>, Error> {
        let mut seed = [0u8; 32];
        rng.fill_bytes(&mut seed);
        let flat = |xs: &Vec<Scalar>| xs.iter().flat_map(|s| s.as_bytes().iter().cloned()).collect::<Vec<u8>>();
        let (a_l, a_r, a_o, v, bl) = (flat(&self.a_l), flat(&self.a_r), flat(&self.a_o), flat(&self.v), flat(&self.v_blinding));
        let cs = self.rec.as_cs(self.v.len());
        let mut proof = vec![0u8; 1 + 32 * (14 + 2 * 32 + 2)];
        let mut len = proof.len();
        check(unsafe {
            sys::bbp_r1cs_prove(self.ctx.raw(), self.transcript.raw, &cs, a_l.as_ptr(), a_r.as_ptr(), a_o.as_ptr(), v.as_ptr(), bl.as_ptr(), seed.as_ptr(),
                                std::ptr::null_mut(), proof.as_mut_ptr(), &mut len)
        })?;
        proof.truncate(len);
        Ok(proof)
    }
}
impl<'a> ConstraintSystem for Prover<'a> {
    fn multiply(&mut self, mut left: LinearCombination, mut right: LinearCombination) -> (Variable, Variable, Variable) {
        let (l, r) = (self.eval(&left), self.eval(&right));
        let i = self.a_l.len();
        self.a_l.push(l);
        self.a_r.push(r);
        self.a_o.push(l * r);
        self.rec.n_multipliers += 1;
        let (lv, rv, ov) = (Variable::MultiplierLeft(i), Variable::MultiplierRight(i), Variable::MultiplierOutput(i));
        left.terms.push((lv, -Scalar::one()));
        right.terms.push((rv, -Scalar::one()));
        self.rec.push(&left);
        self.rec.push(&right);
        (lv, rv, ov)
    }
    fn constrain(&mut self, lc: LinearCombination) {
        self.rec.push(&lc);
    }
}

/// bulletproofs::r1cs::Verifier — `src/blindbid/verify.rs:51-88`.
pub struct Verifier<'a> {
    ctx: &'a Context,
    transcript: &'a mut Transcript,
    rec: Recorded,
    commitments: Vec<u8>,
}
impl<'a> Verifier<'a> {
    pub fn new(ctx: &'a Context, transcript: &'a mut Transcript) -> Self {
        Verifier { ctx, transcript, rec: Recorded::new(), commitments: vec![] }
    }
    /// Verifier::commit (`verify.rs:57`)
    pub fn commit(&mut self, commitment: CompressedRistretto) -> Variable {
        let i = self.commitments.len() / 32;
        self.commitments.extend_from_slice(commitment.as_bytes());
        Variable::Committed(i)
    }
    /// Verifier::verify (`verify.rs:88`): `Ok(())` = accept; `Error::{Format, Verification, InvalidGeneratorsLength}` as upstream.
    pub fn verify<R: RngCore>(self, proof: &[u8], rng: &mut R) -> Result<(), Error> {
        let mut seed = [0u8; 32];
        rng.fill_bytes(&mut seed);
        let cs = self.rec.as_cs(self.commitments.len() / 32);
        check(unsafe { sys::bbp_r1cs_verify(self.ctx.raw(), self.transcript.raw, &cs, proof.as_ptr(), proof.len(), self.commitments.as_ptr(), seed.as_ptr()) })
    }
}
impl<'a> ConstraintSystem for Verifier<'a> {
    fn multiply(&mut self, mut left: LinearCombination, mut right: LinearCombination) -> (Variable, Variable, Variable) {
        let i = self.rec.n_multipliers;
        self.rec.n_multipliers += 1;
        let (lv, rv, ov) = (Variable::MultiplierLeft(i), Variable::MultiplierRight(i), Variable::MultiplierOutput(i));
        left.terms.push((lv, -Scalar::one()));
        right.terms.push((rv, -Scalar::one()));
        self.rec.push(&left);
        self.rec.push(&right);
        (lv, rv, ov)
    }
    fn constrain(&mut self, lc: LinearCombination) {
        self.rec.push(&lc);
    }
}

/// InnerProductProof::create with Q = w * B over the first n resident generators (`bbp_ipp_create`).
pub fn ipp_create(ctx: &Context, transcript: &mut Transcript, w: &Scalar, g_factors: &[Scalar], h_factors: &[Scalar], a: &[Scalar], b: &[Scalar]) -> Result<Vec<u8>, Error> {
    let n = a.len();
    let flat = |xs: &[Scalar]| xs.iter().flat_map(|s| s.as_bytes().iter().cloned()).collect::<Vec<u8>>();
    let (gf, hf, av, bv) = (flat(g_factors), flat(h_factors), flat(a), flat(b));
    let mut out = vec![0u8; 32 * (2 * 32 + 2)];
    let mut len = out.len();
    check(unsafe { sys::bbp_ipp_create(ctx.raw(), transcript.raw, w.as_bytes().as_ptr(), gf.as_ptr(), hf.as_ptr(), av.as_ptr(), bv.as_ptr(), n, out.as_mut_ptr(), &mut len) })?;
    out.truncate(len);
    Ok(out)
}
