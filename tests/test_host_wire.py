"""CPU: the TLV request / reply codec of the outer boundary (csrc/wire.h through the C ABI, no GPU). The byte-level framing is
UNPINNED (dusk-tlv's source is absent); what is pinned here is the STRUCTURE the reference's readers / writers impose
(src/blindbid/proof.rs:97-184, verify.rs:91-128, futures/main.rs:64-110) against an independent Python restatement of the
same recollected framing, and the error behaviour (what yields a reply, what yields none)."""
import os
import struct
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader  # noqa: E402

capi = bbp_loader.load().capi


def tlv(payload):
    n = len(payload)
    if n < 1 << 8:
        return b"\x01" + struct.pack("<B", n) + payload
    if n < 1 << 16:
        return b"\x02" + struct.pack("<H", n) + payload
    return b"\x04" + struct.pack("<I", n) + payload


def tlv_list(items):
    return tlv(b"".join(tlv(x) for x in items))


def sc(i):
    return (i * 0x9e3779b97f4a7c15 % (2**252)).to_bytes(32, "little")


def test_prove_request_layout_and_parse():
    scalars = b"".join(sc(i + 1) for i in range(7))
    pub = b"".join(sc(100 + i) for i in range(5))
    frame = capi.wire_prove_request(scalars, pub, 3)
    want = tlv(b"\x01" + b"".join(tlv(scalars[32 * i:32 * i + 32]) for i in range(7)) + tlv_list([pub[32 * i:32 * i + 32] for i in range(5)]) +
               tlv(struct.pack("<Q", 3)))
    assert frame == want
    st, hdr, pl = capi.wire_frame(frame)
    assert (st, hdr + pl) == (1, len(frame))
    assert capi.wire_frame(frame[:-1])[0] == 0 and capi.wire_frame(frame[:1])[0] == 0
    assert capi.wire_frame(b"\x03" + frame[1:])[0] < 0                 # not a length width
    op, h = capi.wire_parse(frame[hdr:])
    assert op == 1
    capi.wire_free(h)
    # unknown opcode -> 0 (nothing is written); malformed prove bodies -> error (nothing is written)
    assert capi.wire_parse(b"\x07" + frame[hdr + 1:])[0] == 0
    assert capi.wire_parse(frame[hdr:-3])[0] < 0
    short_item = tlv(b"\x01" + b"".join(tlv(scalars[32 * i:32 * i + 32]) for i in range(7)) + tlv_list([pub[:31]]) + tlv(struct.pack("<Q", 0)))
    assert capi.wire_parse(short_item[capi.wire_frame(short_item)[1]:])[0] < 0


def test_proof_blob_roundtrip_and_verify_request():
    proof = bytes(range(256)) * 4 + b"\x09" * 97            # 1121 bytes: needs the 2-byte length form
    comm = b"".join(sc(i + 7) for i in range(4))
    tc = b"".join(sc(i + 70) for i in range(8))
    blob = capi.wire_proof_blob(proof, comm, tc)
    assert blob == tlv(proof) + tlv_list([comm[32 * i:32 * i + 32] for i in range(4)]) + tlv_list([tc[32 * i:32 * i + 32] for i in range(8)])
    assert capi.wire_decode_proof_blob(blob) == (proof, comm, tc)
    with pytest.raises(capi.BbpError):
        capi.wire_decode_proof_blob(blob[:-1])
    bad = tlv(proof) + tlv_list([comm[:31]]) + tlv_list([])
    with pytest.raises(capi.BbpError):
        capi.wire_decode_proof_blob(bad)
    pub = b"".join(sc(200 + i) for i in range(8))
    frame = capi.wire_verify_request(blob, sc(1), sc(2), sc(3), pub)
    assert frame == tlv(b"\x02" + tlv(blob) + tlv(sc(1)) + tlv(sc(2)) + tlv(sc(3)) + tlv_list([pub[32 * i:32 * i + 32] for i in range(8)]))
    st, hdr, pl = capi.wire_frame(frame)
    op, h = capi.wire_parse(frame[hdr:])
    assert op == 2
    capi.wire_free(h)
    # a verify body the reference's readers reject still parses to a request: it is ANSWERED (0x00), not dropped
    op, h = capi.wire_parse(frame[hdr:-5])
    assert op == 2
    capi.wire_free(h)


def test_parsers_survive_random_and_mutated_input():
    """robustness of the outer boundary's parsers on the CPU: random bytes, truncations and single-byte mutations of valid
    frames never crash, never read out of bounds (the process would die under the allocator's guard pages sooner or later)
    and always come back with a definite answer"""
    import random
    rnd = random.Random(1234)
    scalars = b"".join(sc(i + 1) for i in range(7))
    pub = b"".join(sc(100 + i) for i in range(6))
    blob = capi.wire_proof_blob(bytes(range(256)) * 4 + b"\x09" * 97, b"".join(sc(i + 7) for i in range(4)), b"".join(sc(i + 70) for i in range(6)))
    valid = [capi.wire_prove_request(scalars, pub, 2), capi.wire_verify_request(blob, sc(1), sc(2), sc(3), pub)]
    seen = set()
    for frame in valid:
        st, hdr, pl = capi.wire_frame(frame)
        payload = frame[hdr:]
        for cut in list(range(0, min(len(payload), 80))) + [len(payload) - k for k in range(1, 40)]:
            op, h = capi.wire_parse(payload[:cut]) if cut > 0 else (capi.BBP_ERR_INPUT, None)
            seen.add(op)
            if h is not None and op > 0:
                capi.wire_free(h)
        for _ in range(400):
            m = bytearray(payload)
            for _ in range(rnd.randrange(1, 4)):
                m[rnd.randrange(len(m))] = rnd.randrange(256)
            op, h = capi.wire_parse(bytes(m))
            seen.add(op)
            if op > 0:
                capi.wire_free(h)
    for _ in range(300):
        junk = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 200)))
        st, hdr, pl = capi.wire_frame(junk)
        assert st in (0, 1) or st < 0
        op, h = capi.wire_parse(junk)
        seen.add(op)
        if op > 0:
            capi.wire_free(h)
    assert {1, 2, 0} <= seen and any(x < 0 for x in seen)
