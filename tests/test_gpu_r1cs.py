"""The generic bulletproofs surface (bbp_r1cs_prove / bbp_r1cs_verify / bbp_ipp_create / bbp_transcript_*) against the
oracle's generic prover / verifier (oracle/r1cs.h, oracle/ipp.h) on circuits recorded through a ConstraintSystem-shaped
recorder (tests/r1cs_util.py): proof bytes, commitments, verdicts, error classes and the transcript state left behind."""
import hashlib

import pytest

import orc
from orc import L_ORDER, from_le, le
from r1cs_util import LC, Recorder, example_circuit, random_circuit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_pkg():
    from gpu_util import pkg
    return pkg()


@pytest.fixture(scope="module")
def be():
    from gpu_util import backend
    return backend()


@pytest.fixture(scope="module")
def capi(gpu_pkg):
    return gpu_pkg.capi


def blindings(tag, m):
    return b"".join(le(from_le(hashlib.sha512(b"%s-%d" % (tag, i)).digest()) % L_ORDER) for i in range(m))


@pytest.mark.parametrize("seed,n_extra", [(1, 0), (2, 1), (3, 5), (4, 30), (5, 200)])
def test_generic_prove_matches_oracle_and_verifies(be, capi, seed, n_extra):
    cs = example_circuit(seed, n_extra)
    flat = cs.flatten()
    aL, aR, aO, v = cs.witness()
    bl = blindings(b"bl%d" % seed, flat["m"])
    rng = hashlib.sha256(b"rng%d" % seed).digest()
    label = b"generic circuit %d" % seed
    rc, oproof, oV, oafter = orc.r1cs_prove_flat(label, 2048, flat, aL, aR, aO, v, bl, rng)
    assert rc == 0
    tr = capi.Transcript(label)
    st, proof, V = be.r1cs_prove(tr, flat, aL, aR, aO, v, bl, rng)
    assert st == 0
    assert V == oV and proof == oproof
    assert tr.challenge_bytes(b"after", 32) == oafter                    # the transcript is advanced like Prover's &mut borrow
    assert orc.r1cs_verify_flat(label, 2048, flat, proof, V, rng)[0] == 0
    # verification: accept, same final transcript state as the oracle's verifier
    vrng = hashlib.sha256(b"vrng%d" % seed).digest()
    orc_rc, vafter = orc.r1cs_verify_flat(label, 2048, flat, proof, V, vrng)
    tv = capi.Transcript(label)
    assert be.r1cs_verify(tv, flat, proof, V, vrng) == 0 == orc_rc
    assert tv.challenge_bytes(b"after", 32) == vafter
    # wrong label, wrong commitment, a changed coefficient, a changed constant, a bumped proof scalar: all rejected, same class
    def both(lbl, f, p, vv):
        g = be.r1cs_verify(capi.Transcript(lbl), f, p, vv, vrng)
        o = orc.r1cs_verify_flat(lbl, 2048, f, p, vv, vrng)[0]
        assert g == o != 0, (g, o)
    both(label + b"x", flat, proof, V)
    both(label, flat, proof, V[32:64] + V[:32] + V[64:])
    f2 = dict(flat); c = bytearray(flat["term_coeff"]); c[32 * 2:32 * 3] = le((from_le(bytes(c[32 * 2:32 * 3])) + 1) % L_ORDER); f2["term_coeff"] = bytes(c)
    both(label, f2, proof, V)
    p2 = bytearray(proof); p2[-1] ^= 1
    both(label, flat, bytes(p2), V)
    both(label, flat, proof[:-32], V)
    both(label, flat, b"\x07" + proof[1:], V)


def test_unsatisfied_circuit_proves_but_does_not_verify(be, capi):
    """the prover does not check its witness: same bytes as the oracle, rejected by both verifiers"""
    cs = example_circuit(9, 3)
    cs.aO[2] = (cs.aO[2] + 1) % L_ORDER
    flat = cs.flatten()
    aL, aR, aO, v = cs.witness()
    bl, rng = blindings(b"u", flat["m"]), bytes(32)
    rc, oproof, oV, _ = orc.r1cs_prove_flat(b"unsat", 2048, flat, aL, aR, aO, v, bl, rng)
    st, proof, V = be.r1cs_prove(capi.Transcript(b"unsat"), flat, aL, aR, aO, v, bl, rng)
    assert rc == 0 == st and (proof, V) == (oproof, oV)
    assert be.r1cs_verify(capi.Transcript(b"unsat"), flat, proof, V, rng) == orc.r1cs_verify_flat(b"unsat", 2048, flat, proof, V, rng)[0] == -3


def test_circuit_without_commitments_and_capacity_limit(be, capi, gpu_pkg):
    cs = Recorder([])
    a, b = LC.const(6), LC.const(7)
    _, _, o = cs.multiply(a, b)
    cs.constrain(o - LC.const(42))
    flat = cs.flatten()
    aL, aR, aO, v = cs.witness()
    rc, oproof, _, _ = orc.r1cs_prove_flat(b"tiny", 2048, flat, aL, aR, aO, v, b"", bytes(32))
    st, proof, V = be.r1cs_prove(capi.Transcript(b"tiny"), flat, aL, aR, aO, v, b"", bytes(32))
    assert rc == 0 == st and proof == oproof and V == b""
    assert be.r1cs_verify(capi.Transcript(b"tiny"), flat, proof, b"", bytes(32)) == 0
    small = gpu_pkg.Backend(device=0, gens_capacity=4, party_capacity=1)
    big = example_circuit(11, 10)                  # 12 multipliers -> padded 16 > 4
    f = big.flatten()
    aL, aR, aO, v = big.witness()
    st, _, _ = small.r1cs_prove(capi.Transcript(b"cap"), f, aL, aR, aO, v, blindings(b"c", 3), bytes(32))
    assert st == -1 == orc.r1cs_prove_flat(b"cap", 4, f, aL, aR, aO, v, blindings(b"c", 3), bytes(32))[0]
    # malformed circuits are refused, not proven
    bad = dict(flat); bad["term_var"] = [(1 << 28) | 5] + flat["term_var"][1:]
    assert be.r1cs_prove(capi.Transcript(b"tiny"), bad, aL, aR, aO, v, b"", bytes(32))[0] == gpu_pkg.capi.BBP_ERR_FORMAT


def test_transcript_matches_merlin_vector(capi):
    """Merlin's published test vector through the C ABI transcript"""
    t = capi.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    c = t.clone()
    c.append_u64(b"n", 7)
    assert c.challenge_bytes(b"x", 16) != t.challenge_bytes(b"x", 16)


@pytest.mark.parametrize("n", [1, 2, 8, 64, 1024])
def test_standalone_ipp_matches_oracle(be, capi, n):
    sc = lambda tag: orc.random_scalars(hash((tag, n)) & 0xffff, n)
    a, b, gf, hf = sc("a"), sc("b"), sc("g"), sc("h")
    w = orc.random_scalars(77 + n, 1)
    want, after = orc.ipp_create(b"ipp test", w, gf, hf, a, b)
    tr = capi.Transcript(b"ipp test")
    got = be.ipp_create(tr, w, gf, hf, a, b)
    assert got == want
    assert tr.challenge_bytes(b"after", 32) == after


def test_generic_circuit_beyond_blindbid_size(gpu_pkg, capi):
    """2502 multipliers -> padded n = 4096 (lg n = 12, more than any blind-bid circuit): proof bytes == oracle, verifies"""
    be = gpu_pkg.Backend(device=0, gens_capacity=4096, party_capacity=1)
    cs = example_circuit(77, 2500)
    flat = cs.flatten()
    assert capi.cs_shape(flat)[1][:4] == [2502, 5007, 3, 4096]
    aL, aR, aO, v = cs.witness()
    bl, rng = blindings(b"big", 3), hashlib.sha256(b"big").digest()
    rc, oproof, oV, oafter = orc.r1cs_prove_flat(b"big circuit", 4096, flat, aL, aR, aO, v, bl, rng)
    tr = capi.Transcript(b"big circuit")
    st, proof, V = be.r1cs_prove(tr, flat, aL, aR, aO, v, bl, rng)
    assert rc == 0 == st and (proof, V) == (oproof, oV) and tr.challenge_bytes(b"after", 32) == oafter
    assert len(proof) == 1 + 32 * 11 + 64 * 12 + 64
    assert be.r1cs_verify(capi.Transcript(b"big circuit"), flat, proof, V, rng) == 0
    bad = bytearray(proof); bad[500] ^= 2
    assert be.r1cs_verify(capi.Transcript(b"big circuit"), flat, bytes(bad), V, rng) == orc.r1cs_verify_flat(b"big circuit", 4096, flat, bytes(bad), V, rng)[0] != 0
    be.close()


@pytest.mark.parametrize("seed,n_mul,n_commit,n_free", [(1, 1, 0, 0), (2, 3, 1, 2), (3, 17, 4, 9), (4, 64, 2, 30), (5, 200, 7, 100), (6, 333, 0, 5)])
def test_random_circuits_match_oracle(be, capi, seed, n_mul, n_commit, n_free):
    """random satisfiable circuits (random / small / +-1 coefficients, constants, every variable kind on both sides): proof
    bytes and commitments equal the oracle's, both verifiers accept, and a wrong witness is rejected by both"""
    cs = random_circuit(seed, n_mul, n_commit, n_free)
    flat = cs.flatten()
    aL, aR, aO, v = cs.witness()
    bl = blindings(b"rc%d" % seed, flat["m"])
    rng = hashlib.sha256(b"rc%d" % seed).digest()
    label = b"random circuit"
    rc, oproof, oV, oafter = orc.r1cs_prove_flat(label, 2048, flat, aL, aR, aO, v, bl, rng)
    tr = capi.Transcript(label)
    st, proof, V = be.r1cs_prove(tr, flat, aL, aR, aO, v, bl, rng)
    assert rc == 0 == st and (proof, V) == (oproof, oV) and tr.challenge_bytes(b"after", 32) == oafter
    assert be.r1cs_verify(capi.Transcript(label), flat, proof, V, rng) == 0 == orc.r1cs_verify_flat(label, 2048, flat, proof, V, rng)[0]
    if n_mul > 1:
        cs.aL[1] = (cs.aL[1] + 1) % L_ORDER
        aL2 = cs.witness()[0]
        st, p2, V2 = be.r1cs_prove(capi.Transcript(label), flat, aL2, aR, aO, v, bl, rng)
        assert st == 0 and be.r1cs_verify(capi.Transcript(label), flat, p2, V2, rng) == orc.r1cs_verify_flat(label, 2048, flat, p2, V2, rng)[0] == -3
