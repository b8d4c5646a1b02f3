"""Protocol-level checks of the CPU oracle: circuit shape (n1 = 1442+3L, q = 2887+9L, m = 4+L), blind-bid
prove -> verify, single-field mutations reject, format errors, range proofs. The reference holds no tests or vectors
for this path, so proof bytes are "parity unpinned"; these are self-consistency + rejection tests."""
import ctypes
import hashlib

import pytest

import orc
from orc import L_ORDER, from_le, le

lib = orc.lib()
RNG = hashlib.sha256(b"external rng bytes").digest()


@pytest.mark.parametrize("L", [1, 2, 8, 64])
def test_circuit_shape(L):
    out = (ctypes.c_size_t * 3)()
    lib.orc_blindbid_shape(ctypes.c_size_t(L), out)
    assert list(out) == [1442 + 3 * L, 2887 + 9 * L, 4 + L]


@pytest.fixture(scope="module")
def proven():
    bid = orc.make_bid(7, 8, 3)
    rc, proof, comm, tc = orc.blindbid_prove(bid, orc.bid_blindings(7, 8), RNG)
    assert rc == 0
    return bid, proof, comm, tc


def test_prove_verify_accepts(proven):
    bid, proof, comm, tc = proven
    assert len(proof) == 1 + 32 * (3 + 5 + 3 + 2 * 11 + 2) == 1121
    assert orc.blindbid_verify(proof, comm, tc, bid["q"], bid["z_img"], bid["seed"], bid["pub_list"], RNG) == 0
    # a different verifier rng must not change the verdict
    assert orc.blindbid_verify(proof, comm, tc, bid["q"], bid["z_img"], bid["seed"], bid["pub_list"], bytes(32), threads=4) == 0


def test_proof_is_deterministic_and_rng_sensitive(proven):
    bid, proof, comm, tc = proven
    rc, proof2, comm2, tc2 = orc.blindbid_prove(bid, orc.bid_blindings(7, 8), RNG)
    assert (proof2, comm2, tc2) == (proof, comm, tc)
    rc, proof3, comm3, _ = orc.blindbid_prove(bid, orc.bid_blindings(7, 8), bytes(32))
    assert comm3 == comm and proof3 != proof
    # legacy (unversioned, 14-point) layout carries the same elements
    rc, legacy, _, _ = orc.blindbid_prove(bid, orc.bid_blindings(7, 8), RNG, versioned=0)
    assert len(legacy) == 1216 and legacy[:96] == proof[1:97] and legacy[96:192] == bytes(96) and legacy[192:] == proof[97:]
    assert orc.blindbid_verify(legacy, comm, tc, bid["q"], bid["z_img"], bid["seed"], bid["pub_list"], RNG, versioned=0) == 0


def test_mutations_reject(proven):
    bid, proof, comm, tc = proven
    args = dict(score=bid["q"], z_img=bid["z_img"], seed=bid["seed"], pub_list=bid["pub_list"])

    def verify(p=proof, c=comm, t=tc, **kw):
        a = dict(args)
        a.update(kw)
        return orc.blindbid_verify(p, c, t, a["score"], a["z_img"], a["seed"], a["pub_list"], RNG)

    assert verify() == 0
    wrong = le((from_le(bid["q"]) + 1) % L_ORDER)
    assert verify(score=wrong) == -3
    assert verify(z_img=wrong) == -3
    assert verify(seed=wrong) == -3
    pl = bytearray(bid["pub_list"])
    pl[32 * 3] ^= 1
    assert verify(pub_list=bytes(pl)) == -3
    # every 32-byte element of the proof: flip one bit -> reject (format error for non-canonical scalars allowed)
    for off in range(1, len(proof), 32):
        bad = bytearray(proof)
        bad[off + 1] ^= 0x04
        assert verify(p=bytes(bad)) in (-2, -3), off
    # swap two commitments / tamper a toggle commitment
    c2 = comm[32:64] + comm[0:32] + comm[64:]
    assert verify(c=c2) == -3
    t2 = bytearray(tc)
    t2[5] ^= 1
    assert verify(t=bytes(t2)) == -3
    # identity in a validated slot (A_I1) and a non-canonical scalar (t_x = l)
    bad = bytearray(proof)
    bad[1:33] = bytes(32)
    assert verify(p=bytes(bad)) == -3
    bad = bytearray(proof)
    off = 1 + 32 * 8
    bad[off:off + 32] = le(L_ORDER)
    assert verify(p=bytes(bad)) == -2
    # truncated / odd-length proofs
    assert verify(p=proof[:-32]) in (-2, -3)
    assert verify(p=proof[:-1]) == -2
    assert verify(p=b"\x02" + proof[1:]) == -2


def test_wrong_witness_does_not_verify():
    bid = orc.make_bid(9, 4, 1)
    bid = dict(bid)
    bid["toggle"] = 2  # toggle points at an item that is not x
    rc, proof, comm, tc = orc.blindbid_prove(bid, orc.bid_blindings(9, 4), RNG)
    assert rc == 0  # the prover does not check satisfiability
    assert orc.blindbid_verify(proof, comm, tc, bid["q"], bid["z_img"], bid["seed"], bid["pub_list"], RNG) == -3


@pytest.mark.parametrize("L,toggle", [(1, 0), (2, 1), (64, 63)])
def test_other_list_sizes(L, toggle):
    bid = orc.make_bid(100 + L, L, toggle)
    rc, proof, comm, tc = orc.blindbid_prove(bid, orc.bid_blindings(100 + L, L), RNG)
    assert rc == 0 and len(tc) == 32 * L
    assert orc.blindbid_verify(proof, comm, tc, bid["q"], bid["z_img"], bid["seed"], bid["pub_list"], RNG, threads=4) == 0


def test_generator_capacity_limit():
    # L = 203 needs 2051 multipliers > gens capacity 2048 => InvalidGeneratorsLength
    bid = orc.make_bid(5, 203, 0)
    rc, _, _, _ = orc.blindbid_prove(bid, orc.bid_blindings(5, 203), RNG)
    assert rc == -1


def _range_prove(values, nbits, seed=b"\x07" * 32):
    m = len(values)
    vals = (ctypes.c_uint64 * m)(*values)
    bl = b"".join(le(from_le(hashlib.shake_256(b"rp-bl" + bytes([i])).digest(64)) % L_ORDER) for i in range(m))
    proof = ctypes.create_string_buffer(4096)
    plen = ctypes.c_size_t(4096)
    V = ctypes.create_string_buffer(32 * m)
    rc = lib.orc_rangeproof_prove(vals, bl, ctypes.c_size_t(m), ctypes.c_size_t(nbits), seed, proof, ctypes.byref(plen), V)
    return rc, proof.raw[:plen.value], V.raw


def test_rangeproof_small():
    rc, proof, V = _range_prove([0, 255, 17, 128], 8)
    assert rc == 0 and len(proof) == 32 * (7 + 2 * 5 + 2)
    assert lib.orc_rangeproof_verify(proof, ctypes.c_size_t(len(proof)), V, ctypes.c_size_t(4), ctypes.c_size_t(8), RNG, 1) == 0
    bad = bytearray(proof)
    bad[40] ^= 2
    assert lib.orc_rangeproof_verify(bytes(bad), ctypes.c_size_t(len(proof)), V, ctypes.c_size_t(4), ctypes.c_size_t(8), RNG, 1) in (-2, -3)
    # out-of-range value: 256 does not fit 8 bits => the proof must not verify
    rc, proof, V = _range_prove([256, 1], 8)
    assert rc == 0
    assert lib.orc_rangeproof_verify(proof, ctypes.c_size_t(len(proof)), V, ctypes.c_size_t(2), ctypes.c_size_t(8), RNG, 1) == -3


@pytest.mark.slow
def test_rangeproof_config5_shape():
    vals = [from_le(hashlib.shake_256(b"rp-v" + bytes([i])).digest(8)) for i in range(64)]
    rc, proof, V = _range_prove(vals, 64)
    assert rc == 0 and len(proof) == 1056
    assert lib.orc_rangeproof_verify(proof, ctypes.c_size_t(len(proof)), V, ctypes.c_size_t(64), ctypes.c_size_t(64), RNG, 4) == 0
