#!/usr/bin/env python3
"""Generates tests/golden/ristretto_libsodium.json from libsodium 1.0.20's ristretto255 API (an independent
implementation of the group the reference gets from curve25519-dalek 1.2.3) and hashlib. Run in the authoring
container only; the JSON is committed so the GPU box needs neither libsodium nor this script.

  python tests/golden/make_golden.py
"""
import ctypes
import hashlib
import json
import os

SODIUM = "/opt/prime-rl/.venv/lib/python3.12/site-packages/pyzmq.libs/libsodium-19479d6d.so.26.2.0"
S = ctypes.CDLL(SODIUM)
assert S.sodium_init() >= 0
L = 2**252 + 27742317777372353535851937790883648493
P = 2**255 - 19


def buf(n=32):
    return ctypes.create_string_buffer(n)


def shake(tag, n):
    return hashlib.shake_256(tag).digest(n)


def from_hash(h):
    o = buf()
    S.crypto_core_ristretto255_from_hash(o, h)
    return o.raw


def smul(s, p):
    """s*p with identity handled (libsodium returns -1 and zeroes the output for the identity)."""
    o = buf()
    S.crypto_scalarmult_ristretto255(o, s, p)
    return o.raw


def smul_base(s):
    o = buf()
    S.crypto_scalarmult_ristretto255_base(o, s)
    return o.raw


def add(a, b):
    o = buf()
    assert S.crypto_core_ristretto255_add(o, a, b) == 0
    return o.raw


def main():
    g = {}
    # multiples of the generator 0..15 (RFC 9496 appendix A.1 lists the same 16 encodings)
    mult = [bytes(32)]
    for i in range(1, 16):
        mult.append(smul_base(i.to_bytes(32, "little")))
    g["basepoint_multiples"] = [m.hex() for m in mult]
    # from_hash / from_uniform_bytes
    fh = []
    for i in range(64):
        h = shake(b"golden-from-hash" + bytes([i]), 64)
        fh.append([h.hex(), from_hash(h).hex()])
    # edge inputs: all zero, all ones, high bits set
    for h in [bytes(64), b"\xff" * 64, b"\x01" + bytes(63), bytes(32) + b"\xff" * 32]:
        fh.append([h.hex(), from_hash(h).hex()])
    g["from_hash"] = fh
    # scalar * point, point + point
    sm, ad = [], []
    for i in range(48):
        st = shake(b"golden-smul" + bytes([i]), 64 + 64 + 64)
        p = from_hash(st[0:64])
        q = from_hash(st[64:128])
        s = (int.from_bytes(st[128:192], "little") % L).to_bytes(32, "little")
        sm.append([s.hex(), p.hex(), smul(s, p).hex()])
        ad.append([p.hex(), q.hex(), add(p, q).hex()])
    # special scalars
    p = from_hash(shake(b"golden-special", 64))
    for sv in [0, 1, 2, L - 1, L - 2, 2**252, 2**128 - 1]:
        s = sv.to_bytes(32, "little")
        sm.append([s.hex(), p.hex(), smul(s, p).hex()])
    g["scalarmult"] = sm
    g["add"] = ad
    # small multiscalar products (sum over libsodium scalarmult + add)
    msms = []
    for n in [1, 2, 3, 17, 64, 200, 513]:
        st = shake(b"golden-msm" + n.to_bytes(4, "little"), 128 * n)
        pts, scs = [], []
        acc = None
        for i in range(n):
            pt = from_hash(st[128 * i:128 * i + 64])
            s = (int.from_bytes(st[128 * i + 64:128 * i + 128], "little") % L).to_bytes(32, "little")
            pts.append(pt)
            scs.append(s)
            t = smul(s, pt)
            acc = t if acc is None else add(acc, t)
        msms.append({"n": n, "seed_tag": "golden-msm", "result": acc.hex()})
    g["msm"] = msms
    # validity of encodings: crafted + random strings
    enc = []
    crafted = [bytes(32), b"\x01" + bytes(31), b"\xff" * 32, b"\xed" + b"\xff" * 30 + b"\x7f", b"\xec" + b"\xff" * 30 + b"\x7f",
               b"\xee" + b"\xff" * 30 + b"\x7f", bytes(31) + b"\x80", b"\x02" + bytes(31), b"\x04" + bytes(31)]
    for c in crafted:
        enc.append([c.hex(), int(S.crypto_core_ristretto255_is_valid_point(c))])
    for i in range(300):
        c = bytearray(shake(b"golden-enc" + i.to_bytes(4, "little"), 32))
        c[31] &= 0x7f
        c[0] &= 0xfe
        enc.append([bytes(c).hex(), int(S.crypto_core_ristretto255_is_valid_point(bytes(c)))])
    g["is_valid"] = enc
    # scalar arithmetic
    scl = []
    for i in range(32):
        st = shake(b"golden-sc" + bytes([i]), 128)
        a = (int.from_bytes(st[0:64], "little") % L)
        b = (int.from_bytes(st[64:128], "little") % L)
        o = buf()
        S.crypto_core_ristretto255_scalar_mul(o, a.to_bytes(32, "little"), b.to_bytes(32, "little"))
        mul = o.raw
        S.crypto_core_ristretto255_scalar_invert(o, a.to_bytes(32, "little"))
        inv = o.raw
        S.crypto_core_ristretto255_scalar_reduce(o, st[0:64])
        red = o.raw
        assert int.from_bytes(mul, "little") == a * b % L and int.from_bytes(red, "little") == a
        assert int.from_bytes(inv, "little") == pow(a, L - 2, L)
        scl.append([st[0:64].hex(), st[64:128].hex(), red.hex(), mul.hex(), inv.hex()])
    g["scalars"] = scl
    # Pedersen blinding generator: from_hash(SHA3-512(compress(B)))
    g["B_blinding"] = from_hash(hashlib.sha3_512(mult[1]).digest()).hex()
    # BulletproofGens chain samples (SHAKE256("GeneratorsChain" || label || LE32(party)) in 64-byte blocks)
    chains = {}
    for label, party in [(b"G", 0), (b"H", 0), (b"G", 1), (b"H", 63)]:
        st = shake(b"GeneratorsChain" + label + party.to_bytes(4, "little"), 64 * 2048)
        pts = [from_hash(st[64 * i:64 * i + 64]) for i in range(2048)]
        chains["%s%d" % (label.decode(), party)] = {
            "first": [p.hex() for p in pts[:4]], "last": pts[2047].hex(),
            "sha256_of_all": hashlib.sha256(b"".join(pts)).hexdigest()}
    g["gens"] = chains
    # MiMC round constants (src/blindbid/mod.rs:7-24 of the reference): SHA-512 chain, wide-reduced
    h = hashlib.sha512(b"blind bid").digest()
    cs = []
    for _ in range(90):
        c = (int.from_bytes(h, "little") % L).to_bytes(32, "little")
        cs.append(c)
        h = hashlib.sha512(c).digest()
    g["mimc_constants_first"] = cs[0].hex()
    g["mimc_constants_last"] = cs[89].hex()
    g["mimc_constants_sha512"] = hashlib.sha512(b"".join(cs)).hexdigest()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ristretto_libsodium.json")
    with open(out, "w") as f:
        json.dump(g, f, indent=0)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
