"""GPU parity, layer K1/K2/K6: field, group and codec kernels of libbbp_b200.so against the CPU oracle and the
committed libsodium / RFC 9496 vectors. Everything goes through the C ABI (include/bbp.h)."""
import ctypes
import hashlib
import json
import os

import pytest

import orc
from orc import P, from_le, le

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ristretto_libsodium.json")))


def rnd(tag, i, n):
    return hashlib.shake_256(tag + i.to_bytes(4, "little")).digest(n)


@pytest.fixture(scope="module")
def be():
    from gpu_util import backend
    return backend()


def fe_inputs(n):
    edge = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2**255 - 1, 2**255, 2**255 + 18, 2**256 - 1, 2**256 - 38, 2**256 - 39, 2**128, 2**255 - 20]
    a = [from_le(rnd(b"ga", i, 32)) for i in range(n)]
    b = [from_le(rnd(b"gb", i, 32)) for i in range(n)]
    for i, e in enumerate(edge):
        a[i] = e
        b[len(edge) - 1 - i] = e
    for i, e in enumerate(edge):          # all edge x edge pairs as well
        for j, f in enumerate(edge):
            a.append(e)
            b.append(f)
    return a, b


def test_field_ops_match_bigints(be):
    """fe25519.cuh on raw 256-bit limb inputs (the lazy representation) against Python integers."""
    a, b = fe_inputs(2048)
    ab = b"".join(le(x) for x in a)
    bb = b"".join(le(x) for x in b)
    want = {
        0: [x * y % P for x, y in zip(a, b)],
        1: [(x + y) % P for x, y in zip(a, b)],
        2: [(x - y) % P for x, y in zip(a, b)],
        3: [pow(x % P, P - 2, P) for x in a],
        4: [x * x % P for x in a],
        5: [(-x) % P for x in a],
    }
    for op, exp in want.items():
        got = be.test_fe(ab, bb, op)
        got = [from_le(got[32 * i:32 * i + 32]) for i in range(len(a))]
        bad = [i for i in range(len(a)) if got[i] != exp[i]]
        assert not bad, (op, bad[:5], hex(a[bad[0]]), hex(b[bad[0]]))


def test_basepoint_multiples_rfc9496(be):
    """compress(k*B), k = 0..15, via repeated GPU additions, against RFC 9496 §A.1."""
    mult = GOLD["basepoint_multiples"]
    base = bytes.fromhex(mult[1])
    acc = bytes.fromhex(mult[0])
    for k in range(1, 16):
        acc = be.test_ge(acc, base, 0)
        assert acc.hex() == mult[k], k
    # doubling and mixed add / sub
    assert be.test_ge(bytes.fromhex(mult[4]), base, 1).hex() == mult[8]
    assert be.test_ge(bytes.fromhex(mult[4]), bytes.fromhex(mult[3]), 2).hex() == mult[7]
    assert be.test_ge(bytes.fromhex(mult[4]), bytes.fromhex(mult[3]), 3).hex() == mult[1]
    assert be.test_ge(bytes.fromhex(mult[4]), bytes.fromhex(mult[4]), 3).hex() == mult[0]
    assert be.test_ge(base, bytes.fromhex(mult[9]), 4).hex() == mult[9]


def test_decompress_validity_list(be):
    """RFC 9496 invalid encodings + valid ones: flags must match libsodium's is_valid_point (identity is valid)."""
    encs = [bytes.fromhex(h) for h, _ in GOLD["is_valid"]]
    want = [v for _, v in GOLD["is_valid"]]
    ext, valid = be.decompress(b"".join(encs))
    assert list(valid) == want
    # valid points round trip bit-exactly through GPU decompress -> GPU compress
    good = [e for e, v in zip(encs, want) if v]
    ext, valid = be.decompress(b"".join(good))
    assert be.compress(ext) == b"".join(good)
    # and the extended coordinates agree with the oracle's decompression
    lib = orc.lib()
    out = ctypes.create_string_buffer(128)
    for i, e in enumerate(good[:64]):
        assert lib.orc_ge_decompress_ext(out, e) == 1
        assert out.raw == ext[128 * i:128 * i + 128]


def test_from_uniform_bytes(be):
    """from_uniform_bytes (Elligator x2 + add) against the libsodium from_hash vectors and the oracle."""
    ins = b"".join(bytes.fromhex(h) for h, _ in GOLD["from_hash"])
    want = b"".join(bytes.fromhex(p) for _, p in GOLD["from_hash"])
    assert be.from_uniform_bytes(ins) == want
    n = 1000
    stream = hashlib.shake_256(b"gpu-uniform").digest(64 * n)
    got = be.from_uniform_bytes(stream)
    lib = orc.lib()
    out = ctypes.create_string_buffer(32)
    for i in range(n):
        lib.orc_ge_from_uniform(out, stream[64 * i:64 * i + 64])
        assert out.raw == got[32 * i:32 * i + 32], i


def test_add_and_scalarmult_vectors(be):
    """point additions against the libsodium vectors; compress on non-trivial Z (sums) exercises both compress branches."""
    a = b"".join(bytes.fromhex(x[0]) for x in GOLD["add"])
    b = b"".join(bytes.fromhex(x[1]) for x in GOLD["add"])
    s = b"".join(bytes.fromhex(x[2]) for x in GOLD["add"])
    assert be.test_ge(a, b, 0) == s
    assert be.test_ge(a, b, 2) == s


def test_generators_match_golden_and_oracle(be):
    """Resident generator tables (built once at bbp_init) against §4.3's golden values and the oracle's chain."""
    B, Bb = be.pedersen_gens()
    assert B.hex() == GOLD["basepoint_multiples"][1]
    assert Bb.hex() == GOLD["B_blinding"]
    lib = orc.lib()
    for which in "GH":
        got = be.bulletproof_gens(which, 0, 0, 2048)
        out = ctypes.create_string_buffer(32 * 2048)
        lib.orc_bp_gens(out, ord(which), 0, ctypes.c_size_t(2048))
        assert got == out.raw
        g = GOLD["gens"][which + "0"]
        for i, h in enumerate(g["first"]):
            assert got[32 * i:32 * i + 32].hex() == h
        assert got[32 * 2047:].hex() == g["last"]
