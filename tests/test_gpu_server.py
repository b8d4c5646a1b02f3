"""The outer boundary end to end on the GPU: bbp_wire_execute (decode -> batched prove / verify -> encode) and the
Unix-socket server shell with concurrent clients (one TLV request per connection, replies as src/futures/main.rs:64-110
prescribes: a proof blob, one verdict byte, or nothing at all)."""
import hashlib
import os
import socket
import subprocess
import tempfile
import threading
import time

import pytest

import orc
from orc import L_ORDER, from_le, le

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def be():
    from gpu_util import backend
    return backend()


@pytest.fixture(scope="module")
def capi():
    from gpu_util import pkg
    return pkg().capi


def bid_scalars(b):
    return b"".join(b[k] for k in ("d", "k", "y", "y_inv", "q", "z_img", "seed"))


def payload_of(capi, frame):
    st, hdr, pl = capi.wire_frame(frame)
    assert st == 1 and hdr + pl == len(frame)
    return frame[hdr:]


def test_wire_execute_matches_direct_calls(be, capi):
    L = 4
    bids = [orc.make_bid(300 + i, L, i % L) for i in range(6)]
    seed = hashlib.sha256(b"wire").digest()
    handles = []
    for b in bids:
        op, h = capi.wire_parse(payload_of(capi, capi.wire_prove_request(bid_scalars(b), b["pub_list"], b["toggle"])))
        assert op == 1
        handles.append(h)
    # a non-canonical scalar: the reference's serde rejects it and writes nothing
    bad = dict(bids[0]); bad["k"] = le(L_ORDER + 3)
    op, h = capi.wire_parse(payload_of(capi, capi.wire_prove_request(bid_scalars(bad), bad["pub_list"], 0)))
    assert op == 1
    handles.append(h)
    replies = capi.wire_execute(be, handles, seed)
    for h in handles:
        capi.wire_free(h)
    assert replies[-1] is None
    items = []
    for i, (b, rep) in enumerate(zip(bids, replies[:-1])):
        blob = payload_of(capi, rep)
        proof, comm, tc = capi.wire_decode_proof_blob(blob)
        assert len(proof) == 1121 and len(comm) == 128 and len(tc) == 32 * L
        # same bytes as the direct entry point under the same derived randomness: SHAKE256(seed || LE64(index))
        rnd = hashlib.shake_256(seed + i.to_bytes(8, "little")).digest(64 * (4 + L) + 32)
        bl = b"".join(le(from_le(rnd[64 * k:64 * k + 64]) % L_ORDER) for k in range(4 + L))
        direct = be.blindbid_prove(dict(b, blindings=bl, rng_seed=rnd[-32:]))
        assert direct == (0, proof, comm, tc)
        assert orc.blindbid_verify(proof, comm, tc, b["q"], b["z_img"], b["seed"], b["pub_list"], bytes(32)) == 0
        items.append((b, blob))
    # verification: honest, wrong score, corrupted proof, malformed body
    vh = []
    for b, blob in items[:3]:
        vh.append(capi.wire_parse(payload_of(capi, capi.wire_verify_request(blob, b["q"], b["z_img"], b["seed"], b["pub_list"])))[1])
    b, blob = items[3]
    vh.append(capi.wire_parse(payload_of(capi, capi.wire_verify_request(blob, le(5), b["z_img"], b["seed"], b["pub_list"])))[1])
    b, blob = items[4]
    cb = bytearray(blob); cb[700] ^= 1
    vh.append(capi.wire_parse(payload_of(capi, capi.wire_verify_request(bytes(cb), b["q"], b["z_img"], b["seed"], b["pub_list"])))[1])
    b, blob = items[5]
    op, h = capi.wire_parse(payload_of(capi, capi.wire_verify_request(blob, b["q"], b["z_img"], b["seed"], b["pub_list"]))[:-7])
    assert op == 2
    vh.append(h)
    vr = capi.wire_execute(be, vh, seed)
    for h in vh:
        capi.wire_free(h)
    assert [payload_of(capi, r) for r in vr] == [b"\x01"] * 3 + [b"\x00"] * 3


def _roundtrip(path, frame):
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(60)
    s.connect(path)
    s.sendall(frame)
    out = b""
    while True:
        try:
            chunk = s.recv(65536)
        except ConnectionResetError:      # the server dropped a request it would not answer with bytes still unread
            break
        if not chunk:
            break
        out += chunk
    s.close()
    return out


def test_server_coalesces_concurrent_clients(capi):
    srv = os.path.join(ROOT, "dusk-blindbidproof_b200", "bbp-blindbid-server")
    if not os.path.exists(srv):
        pytest.fail("the server shell was not built (build.sh)")
    path = os.path.join(tempfile.mkdtemp(), "uds")
    proc = subprocess.Popen([srv, "-b", path, "-l", "debug", "--window-us", "20000"], stderr=subprocess.PIPE, text=True)
    try:
        for _ in range(600):
            if os.path.exists(path):
                break
            assert proc.poll() is None, proc.stderr.read()
            time.sleep(0.1)
        assert os.path.exists(path)
        L, n = 8, 24
        bids = [orc.make_bid(400 + i, L, i % L) for i in range(n)]
        frames = [capi.wire_prove_request(bid_scalars(b), b["pub_list"], b["toggle"]) for b in bids]
        replies = [None] * n

        def client(i, fr):
            replies[i] = _roundtrip(path, fr)

        th = [threading.Thread(target=client, args=(i, f)) for i, f in enumerate(frames)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        blobs = []
        for b, rep in zip(bids, replies):
            blob = payload_of(capi, rep)
            proof, comm, tc = capi.wire_decode_proof_blob(blob)
            assert orc.blindbid_verify(proof, comm, tc, b["q"], b["z_img"], b["seed"], b["pub_list"], bytes(32)) == 0 if len(blobs) < 3 else True
            blobs.append(blob)
        # verify through the socket, one of them with the wrong seed; an unknown opcode and a truncated frame get no reply
        vframes = [capi.wire_verify_request(blob, b["q"], b["z_img"], b["seed"], b["pub_list"]) for b, blob in zip(bids, blobs)]
        vframes[5] = capi.wire_verify_request(blobs[5], bids[5]["q"], bids[5]["z_img"], le(1234), bids[5]["pub_list"])
        vreplies = [None] * n
        th = [threading.Thread(target=lambda i=i, f=f: vreplies.__setitem__(i, _roundtrip(path, f))) for i, f in enumerate(vframes)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        want = [b"\x01"] * n
        want[5] = b"\x00"
        assert [payload_of(capi, r) for r in vreplies] == want
        # a slow client: the frame arrives in pieces (tag byte alone, half of the length field, the payload in three parts),
        # which walks the reader's state machine through every partial-read branch
        fr = vframes[0]
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.settimeout(60)
        s.connect(path)
        cuts = [0, 1, 2, 3, 40, len(fr) // 2, len(fr)]
        for a, b in zip(cuts, cuts[1:]):
            s.sendall(fr[a:b])
            time.sleep(0.02)
        got = b""
        while True:
            chunk = s.recv(4096)
            if not chunk:
                break
            got += chunk
        s.close()
        assert payload_of(capi, got) == b"\x01"
        # a client that sends half a frame and hangs up is dropped without a reply and without disturbing the others
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.connect(path)
        s.sendall(fr[:100])
        s.close()
        assert payload_of(capi, _roundtrip(path, vframes[1])) == b"\x01"
        unknown = bytearray(frames[0]); unknown[capi.wire_frame(frames[0])[1]] = 9
        assert _roundtrip(path, bytes(unknown)) == b""
        assert _roundtrip(path, b"\x05garbage") == b""
    finally:
        proc.terminate()
        try:
            err = proc.communicate(timeout=30)[1]
        except subprocess.TimeoutExpired:
            proc.kill()
            err = proc.communicate()[1]
    # the 24 concurrent prove requests were served by far fewer GPU batches than requests
    import re
    m = re.search(r"served (\d+) requests in (\d+) batches", err)
    assert m, err[-2000:]
    assert int(m.group(1)) == 2 * n + 2 and int(m.group(2)) <= 14, err[-2000:]
