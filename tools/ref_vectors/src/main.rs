//! Emits golden vectors produced BY THE REFERENCE's own code path: its gadgets (`dusk_blindbidproof::gadgets`), its constants,
//! its transcript label and generator sizes (`generate_cs_transcript`), its proof blob (`R1CSProof::to_bytes`), driven
//! exactly as `Proof::prove` drives them (src/blindbid/proof.rs:36-91) except that the commitment blindings are explicit
//! and the 32 bytes bulletproofs takes from `thread_rng` come from BBP_REF_RNG32 (see bulletproofs-seeded-rng.patch).
//!
//! usage: bbp-ref-vectors <list length L> <count>  -> JSON on stdout (tests/golden/reference_L<L>.json)
use bulletproofs::r1cs::{Prover, Verifier};
use curve25519_dalek::ristretto::CompressedRistretto;
use curve25519_dalek::scalar::Scalar;
use dusk_blindbidproof::gadgets;
use dusk_blindbidproof::blindbid::generate_cs_transcript;
use sha2::{Digest, Sha512};
use rand::{Rng, SeedableRng};
use rand_chacha::ChaChaRng;

/// `CONSTANTS` is private to the reference (src/blindbid/mod.rs:7-24): the same SHA-512 chain, computed here
fn constants() -> Vec<Scalar> {
    let mut out = Vec::with_capacity(90);
    let mut hash = [0u8; 64];
    hash.copy_from_slice(Sha512::digest(b"blind bid").as_slice());
    for _ in 0..90 {
        let c = Scalar::from_bytes_mod_order_wide(&hash);
        out.push(c);
        hash.copy_from_slice(Sha512::digest(&c.to_bytes()).as_slice());
    }
    out
}

fn mimc(consts: &[Scalar], left: Scalar, right: Scalar) -> Scalar {
    // the permutation gadgets::mimc_gadget constrains (src/gadgets.rs:45-67)
    let mut x = left;
    for c in consts.iter() {
        let a = x + right + c;
        let a2 = a * a;
        let a3 = a2 * a;
        let a4 = a2 * a2;
        x = a4 * a3;
    }
    x + right
}

fn hexs(s: &Scalar) -> String { hex::encode(s.as_bytes()) }

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let l: usize = args.get(1).map(|s| s.parse().unwrap()).unwrap_or(8);
    let count: usize = args.get(2).map(|s| s.parse().unwrap()).unwrap_or(2);
    let consts = constants();
    let mut out = Vec::new();
    for case in 0..count {
        let mut rng = ChaChaRng::seed_from_u64(0xbb9 + 1000 * l as u64 + case as u64);
        let k = Scalar::random(&mut rng);
        let d = Scalar::from(rng.gen::<u64>());
        let seed = Scalar::random(&mut rng);
        let m = mimc(&consts, k, Scalar::zero());
        let x = mimc(&consts, d, m);
        let y = mimc(&consts, seed, x);
        let z_img = mimc(&consts, seed, m);
        let y_inv = y.invert();
        let q = d * y_inv;
        let toggle = case % l;
        let mut pub_list: Vec<Scalar> = (0..l).map(|_| Scalar::random(&mut rng)).collect();
        pub_list[toggle] = x;
        let blindings: Vec<Scalar> = (0..4 + l).map(|_| Scalar::random(&mut rng)).collect();
        let mut rng32 = [0u8; 32];
        rng.fill(&mut rng32);
        std::env::set_var("BBP_REF_RNG32", hex::encode(rng32));

        // ---- Proof::prove, src/blindbid/proof.rs:47-90
        let (pc_gens, bp_gens, mut transcript) = generate_cs_transcript();
        let mut prover = Prover::new(&pc_gens, &mut transcript);
        let mut commitments = Vec::new();
        let mut vars = Vec::new();
        for (v, b) in [d, k, y, y_inv].iter().zip(blindings.iter()) {
            let (c, var) = prover.commit(*v, *b);
            commitments.push(c);
            vars.push(var);
        }
        let mut t_c = Vec::new();
        let mut t_v = Vec::new();
        for i in 0..l {
            let bit = if i == toggle { Scalar::one() } else { Scalar::zero() };
            let (c, var) = prover.commit(bit, blindings[4 + i]);
            t_c.push(c);
            t_v.push(var);
        }
        let items: Vec<_> = pub_list.iter().map(|s| (*s).into()).collect();
        gadgets::proof_gadget(&mut prover, vars[0].into(), vars[1].into(), vars[3].into(), q.into(), z_img.into(), seed.into(), &consts, t_v, items);
        let proof = prover.prove(&bp_gens).expect("prove");
        let proof_bytes = proof.to_bytes();

        // ---- Verify::verify, src/blindbid/verify.rs:47-89
        let (pc_gens, bp_gens, mut transcript) = generate_cs_transcript();
        let mut verifier = Verifier::new(&mut transcript);
        let vars: Vec<_> = commitments.iter().map(|c: &CompressedRistretto| verifier.commit(*c)).collect();
        let t_v: Vec<_> = t_c.iter().map(|c| verifier.commit(*c)).collect();
        let items: Vec<_> = pub_list.iter().map(|s| (*s).into()).collect();
        gadgets::proof_gadget(&mut verifier, vars[0].into(), vars[1].into(), vars[3].into(), q.into(), z_img.into(), seed.into(), &consts, t_v, items);
        let verdict = verifier.verify(&proof, &pc_gens, &bp_gens).is_ok();

        out.push(serde_json::json!({
            "L": l, "toggle": toggle,
            "d": hexs(&d), "k": hexs(&k), "y": hexs(&y), "y_inv": hexs(&y_inv), "q": hexs(&q), "z_img": hexs(&z_img), "seed": hexs(&seed),
            "pub_list": pub_list.iter().map(hexs).collect::<Vec<_>>(),
            "blindings": blindings.iter().map(hexs).collect::<Vec<_>>(),
            "rng_seed": hex::encode(rng32),
            "proof": hex::encode(&proof_bytes), "proof_len": proof_bytes.len(),
            "commitments": commitments.iter().map(|c| hex::encode(c.as_bytes())).collect::<Vec<_>>(),
            "t_c": t_c.iter().map(|c| hex::encode(c.as_bytes())).collect::<Vec<_>>(),
            "verdict": verdict,
        }));
    }
    println!("{}", serde_json::to_string_pretty(&serde_json::json!({
        "source": "dusk-blindbidproof reference + bulletproofs 4a05305 (seeded-rng patch), tools/ref_vectors",
        "vectors": out })).unwrap());
}
