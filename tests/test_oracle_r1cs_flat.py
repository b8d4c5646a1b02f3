"""CPU: the oracle's flattened-circuit entry points (what tests/test_gpu_r1cs.py compares the product with) are
self-consistent: prove -> verify accepts, mutations reject, the standalone IPP has the documented length."""
import hashlib

import orc
from orc import L_ORDER, from_le, le
from r1cs_util import example_circuit, random_circuit


def test_flat_prove_verify_roundtrip():
    cs = example_circuit(21, 4)
    flat = cs.flatten()
    aL, aR, aO, v = cs.witness()
    bl = b"".join(le(from_le(hashlib.sha512(b"b%d" % i).digest()) % L_ORDER) for i in range(flat["m"]))
    rc, proof, V, after = orc.r1cs_prove_flat(b"cpu flat", 64, flat, aL, aR, aO, v, bl, bytes(32))
    assert rc == 0 and len(proof) == 1 + 32 * (8 + 3) + 64 * 3 + 64      # 6 multipliers -> n = 8, lg n = 3
    assert orc.r1cs_verify_flat(b"cpu flat", 64, flat, proof, V, bytes(32))[0] == 0
    assert orc.r1cs_verify_flat(b"cpu flaT", 64, flat, proof, V, bytes(32))[0] == -3
    bad = bytearray(proof); bad[40] ^= 1
    assert orc.r1cs_verify_flat(b"cpu flat", 64, flat, bytes(bad), V, bytes(32))[0] != 0
    assert orc.r1cs_prove_flat(b"cpu flat", 4, flat, aL, aR, aO, v, bl, bytes(32))[0] == -1


def test_ipp_length():
    n = 8
    s = lambda k: orc.random_scalars(k, n)
    out, _ = orc.ipp_create(b"x", orc.random_scalars(9, 1), s(1), s(2), s(3), s(4))
    assert len(out) == 32 * (2 * 3 + 2)


def test_random_circuits_are_satisfiable_for_the_oracle():
    for seed, n_mul, n_commit, n_free in [(2, 3, 1, 2), (3, 17, 4, 9)]:
        cs = random_circuit(seed, n_mul, n_commit, n_free)
        flat = cs.flatten()
        aL, aR, aO, v = cs.witness()
        bl = b"".join(le(from_le(hashlib.sha512(b"r%d" % i).digest()) % L_ORDER) for i in range(flat["m"]))
        rc, proof, V, _ = orc.r1cs_prove_flat(b"rnd", 64, flat, aL, aR, aO, v, bl, bytes(32))
        assert rc == 0 and orc.r1cs_verify_flat(b"rnd", 64, flat, proof, V, bytes(32))[0] == 0
