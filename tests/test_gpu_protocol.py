"""GPU parity, protocol layer: Pedersen commitments, generator-table MSMs, the blind-bid prover (proof bytes identical
to the oracle's under the same blindings / rng seed — BASELINE config 3) and verifier (same verdicts as the oracle on
honest, mutated and malformed proofs — configs 1, 4). Everything goes through the C ABI."""
import hashlib

import pytest

import orc
from orc import L_ORDER, from_le, le

pytestmark = pytest.mark.gpu
ZERO32 = bytes(32)


@pytest.fixture(scope="module")
def be():
    from gpu_util import backend
    return backend()


def seed32(tag):
    return hashlib.sha256(tag.encode()).digest()


def make_case(i, L, toggle=None):
    toggle = (i % L) if toggle is None else toggle
    bid = orc.make_bid(1000 + i, L, toggle)
    bid["blindings"] = orc.bid_blindings(1000 + i, L)
    bid["rng_seed"] = seed32(f"rng{i}")
    return bid


def verify_item(bid, proof, comm, tc, i=0):
    return dict(proof=proof, commitments=comm, t_c=tc, score=bid["q"], z_img=bid["z_img"], seed=bid["seed"], pub_list=bid["pub_list"],
                rng_seed=seed32(f"vrng{i}"))


def oracle_verify(item):
    return orc.blindbid_verify(item["proof"], item["commitments"], item["t_c"], item["score"], item["z_img"], item["seed"], item["pub_list"],
                               item["rng_seed"], threads=4)


def split_proof(p):
    """versioned one-phase layout: 1 + 8 points + 3 scalars + ipp"""
    names = ["A_I1", "A_O1", "S1", "T_1", "T_3", "T_4", "T_5", "T_6", "t_x", "t_x_blinding", "e_blinding"]
    out = {"version": p[:1]}
    pos = 1
    for n in names:
        out[n] = p[pos:pos + 32]
        pos += 32
    k = 0
    while pos + 64 < len(p):
        out[f"L{k}"] = p[pos:pos + 32]
        out[f"R{k}"] = p[pos + 32:pos + 64]
        pos += 64
        k += 1
    out["a"], out["b"] = p[pos:pos + 32], p[pos + 32:pos + 64]
    return out


def test_pedersen_commit(be):
    B, Bb = be.pedersen_gens()
    n = 40
    vals = bytearray(orc.random_scalars(31, n))
    bls = bytearray(orc.random_scalars(32, n))
    vals[0:32] = le(0); bls[32:64] = le(0); vals[64:96] = le(1); bls[64:96] = le(0)
    vals[96:128] = le(L_ORDER - 1); bls[96:128] = le(L_ORDER - 1); vals[128:160] = le(0); bls[128:160] = le(0)
    got = be.pedersen_commit(bytes(vals), bytes(bls))
    for i in range(n):
        want = orc.msm(bytes(vals[32 * i:32 * i + 32]) + bytes(bls[32 * i:32 * i + 32]), B + Bb, algo=0)
        assert got[32 * i:32 * i + 32] == want, i
    assert got[64:96] == B and got[128:160] == ZERO32


def test_msm_over_generator_table(be):
    """fixed-base (window table) MSM slots over [B, B_blinding, G, H] against the oracle over the compressed generators."""
    B, Bb = be.pedersen_gens()
    gens = B + Bb + be.bulletproof_gens("G", 0, 0, 2048) + be.bulletproof_gens("H", 0, 0, 2048)
    slot_len = 4098
    slots = 3
    scs = bytearray(orc.random_scalars(41, slot_len * slots))
    # slot 1: sparse (like A_O1: only B_blinding and part of G), slot 2: a few special scalars
    for i in range(slot_len):
        if i == 0 or i > 1500:
            scs[32 * (slot_len + i):32 * (slot_len + i + 1)] = ZERO32
    for k, s in enumerate([0, 1, L_ORDER - 1, 2**252, 1 << 11, (1 << 11) - 1, 1 << 10, (1 << 10) + 1]):
        scs[32 * (2 * slot_len + k):32 * (2 * slot_len + k + 1)] = le(s)
    scs = bytes(scs)
    got = be.msm_gens(scs, slot_len, slots)
    for k in range(slots):
        assert got[32 * k:32 * k + 32] == orc.msm(scs[32 * slot_len * k:32 * slot_len * (k + 1)], gens, algo=1, threads=4), k
    # short slots (2 columns) = Pedersen commitments through the same engine
    two = orc.random_scalars(42, 2 * 5)
    got = be.msm_gens(two, 2, 5)
    for k in range(5):
        assert got[32 * k:32 * k + 32] == orc.msm(two[64 * k:64 * k + 64], B + Bb, algo=0)


@pytest.mark.parametrize("small_msm", ["latency path", "engine"])
@pytest.mark.parametrize("L", [1, 2, 8])
def test_prove_bytes_identical_to_oracle(be, L, small_msm, monkeypatch):
    """BASELINE config 3: proof bytes, commitments and toggle commitments equal the oracle's under fixed blindings / rng.
    Single requests take the two-launch digit-table MSMs (small_msm.cuh) by default; BBP_SMALL_MSM_MAX=0 sends the same
    request through the bucket engine."""
    if small_msm == "engine":
        monkeypatch.setenv("BBP_SMALL_MSM_MAX", "0")
    bid = make_case(L, L)
    rc, proof, comm, tc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"])
    assert rc == 0
    st, gproof, gcomm, gtc = be.blindbid_prove(bid)
    assert st == 0
    assert gcomm == comm and gtc == tc
    want, got = split_proof(proof), split_proof(gproof)
    for k in want:
        assert got.get(k) == want[k], f"proof field {k} differs"
    assert gproof == proof and len(gproof) == 1121
    # the oracle accepts the GPU proof, the GPU accepts the oracle's proof
    item = verify_item(bid, gproof, gcomm, gtc)
    assert oracle_verify(item) == 0
    assert be.blindbid_verify(item) == 0


def test_prove_batch_mixed_lengths(be):
    """a batch mixing list lengths and toggles: every proof equals the single-proof oracle output"""
    cases = [make_case(10 + i, L) for i, L in enumerate([8, 8, 3, 8, 1, 3, 8, 8])]
    outs = be.blindbid_prove_batch(cases)
    for bid, (st, proof, comm, tc) in zip(cases, outs):
        rc, oproof, ocomm, otc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"])
        assert st == 0 and rc == 0
        assert (proof, comm, tc) == (oproof, ocomm, otc)


@pytest.mark.parametrize("keccak", ["warp", "thread"])
def test_prove_batch_device_rng_matches_oracle(be, keccak, monkeypatch):
    """large batches continue the TranscriptRng on the device (rng_kernels.cuh: one warp per proof, or the older
    one-thread-per-proof kernel) and switch the inner-product argument to its hybrid form (materialised folded bases):
    bytes must still equal the oracle's, and equal what the host-RNG / plain path (single request) produces"""
    cases = [make_case(300 + i, 3) for i in range(12)]
    monkeypatch.setenv("BBP_KECCAK_THREAD", "1" if keccak == "thread" else "0")
    monkeypatch.setenv("BBP_DEVICE_RNG_MIN_BATCH", "8")
    monkeypatch.setenv("BBP_IPP_HYBRID", "2")
    outs = be.blindbid_prove_batch(cases)
    monkeypatch.delenv("BBP_DEVICE_RNG_MIN_BATCH")
    monkeypatch.delenv("BBP_IPP_HYBRID")
    for i, (bid, (st, proof, comm, tc)) in enumerate(zip(cases, outs)):
        assert st == 0
        if i % 4 == 0:
            rc, oproof, ocomm, otc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"])
            assert rc == 0 and (proof, comm, tc) == (oproof, ocomm, otc), i
        if i % 4 == 1:
            assert be.blindbid_prove(bid) == (st, proof, comm, tc)


def test_prove_batch_lanes_match_oracle(be, monkeypatch):
    """batches above the part size (1024; shrunk here) are cut into parts proved concurrently on sibling contexts of the
    same GPU: the bytes must not depend on the cut, and the launch counter must include the sibling contexts"""
    cases = [make_case(500 + i, L) for i, L in enumerate([2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 5, 5, 5, 5])]
    monkeypatch.setenv("BBP_PROVE_PART", "4")
    monkeypatch.setenv("BBP_PROVE_LANES", "3")
    l0 = be.launch_count()
    outs = be.blindbid_prove_batch(cases)
    l1 = be.launch_count()
    monkeypatch.setenv("BBP_PROVE_LANES", "1")
    serial = be.blindbid_prove_batch(cases)
    l2 = be.launch_count()
    assert outs == serial
    assert l1 - l0 == l2 - l1 > 0
    for i in (0, 3, 4, 9, 10, 13):
        bid = cases[i]
        rc, oproof, ocomm, otc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"])
        assert rc == 0 and outs[i] == (0, oproof, ocomm, otc), i


def mutations(bid, proof, comm, tc):
    """(name, item) pairs; each must be rejected. Mirrors SURVEY.md §4.4-3."""
    out = []
    base = verify_item(bid, proof, comm, tc)
    f = split_proof(proof)

    def with_proof(name, p):
        it = dict(base); it["proof"] = p
        out.append((name, it))

    off = 1
    for idx, name in enumerate(["A_I1", "A_O1", "S1", "T_1", "T_3", "T_4", "T_5", "T_6"]):
        p = bytearray(proof)
        p[off + 32 * idx:off + 32 * idx + 32] = f["A_O1"] if name != "A_O1" else f["A_I1"]   # a valid but wrong point
        with_proof("swap " + name, bytes(p))
    p = bytearray(proof); p[1:33] = ZERO32; with_proof("identity A_I1", bytes(p))
    p = bytearray(proof); p[1 + 32 * 3:1 + 32 * 4] = ZERO32; with_proof("identity T_1", bytes(p))
    p = bytearray(proof); p[1] |= 1; with_proof("invalid point encoding A_I1", bytes(p))
    for idx, name in enumerate(["t_x", "t_x_blinding", "e_blinding"]):
        p = bytearray(proof)
        v = (from_le(f[name]) + 1) % L_ORDER
        p[off + 32 * (8 + idx):off + 32 * (9 + idx)] = le(v)
        with_proof("bump " + name, bytes(p))
    p = bytearray(proof); p[off + 32 * 8:off + 32 * 9] = le(L_ORDER); with_proof("non-canonical t_x", bytes(p))
    p = bytearray(proof); p[off + 32 * 11:off + 32 * 12] = f["R0"]; with_proof("swap L0", bytes(p))
    p = bytearray(proof); p[off + 32 * 11:off + 32 * 12] = ZERO32; with_proof("identity L0", bytes(p))
    p = bytearray(proof); p[-32:] = le((from_le(f["b"]) + 1) % L_ORDER); with_proof("bump b", bytes(p))
    p = bytearray(proof); p[-64:-32] = le(2**255 - 1); with_proof("non-canonical a", bytes(p))
    with_proof("truncated", proof[:-32])
    with_proof("one byte short", proof[:-1])
    with_proof("two rounds dropped", proof[:-64 - 128] + proof[-64:])
    with_proof("bad version byte", b"\x02" + proof[1:])
    with_proof("empty", b"")
    it = dict(base); it["score"] = le((from_le(bid["q"]) + 1) % L_ORDER); out.append(("wrong score", it))
    it = dict(base); it["z_img"] = le((from_le(bid["z_img"]) + 1) % L_ORDER); out.append(("wrong z_img", it))
    it = dict(base); it["seed"] = le((from_le(bid["seed"]) + 1) % L_ORDER); out.append(("wrong seed", it))
    L = bid["L"]
    pl = bytearray(bid["pub_list"]); t = bid["toggle"]
    pl[32 * t:32 * t + 32] = le((from_le(pl[32 * t:32 * t + 32]) + 1) % L_ORDER)
    it = dict(base); it["pub_list"] = bytes(pl); out.append(("own item changed in the list", it))
    c = bytearray(comm); c[0:32], c[32:64] = comm[32:64], comm[0:32]
    it = dict(base); it["commitments"] = bytes(c); out.append(("commitments swapped", it))
    c = bytearray(comm); c[0] |= 1
    it = dict(base); it["commitments"] = bytes(c); out.append(("invalid commitment encoding", it))
    it = dict(base); it["commitments"] = comm[:96]; out.append(("three commitments", it))
    it = dict(base); it["t_c"] = b""; out.append(("no toggles", it))
    if L > 1:
        it = dict(base); it["t_c"] = tc[:-32]; out.append(("one toggle commitment missing", it))
        it = dict(base); it["pub_list"] = bid["pub_list"][:-32]; out.append(("list shorter than toggles", it))
    return out


@pytest.mark.parametrize("L,replay", [(2, "host"), (8, "host"), (8, "device"), (3, "device"), (8, "device-thread")])
def test_verify_verdicts_match_oracle(be, L, replay, monkeypatch):
    """accept / reject parity with the oracle, including the error class, over honest and mutated inputs; with the
    Fiat-Shamir replay on the host threads (small batches) and on the device (large batches: one warp per request, or
    the older one-thread-per-request kernel)"""
    monkeypatch.setenv("BBP_DEVICE_TRANSCRIPT_MIN_BATCH", "1" if replay.startswith("device") else "1000000")
    monkeypatch.setenv("BBP_KECCAK_THREAD", "1" if replay == "device-thread" else "0")
    bid = make_case(50 + L, L)
    rc, proof, comm, tc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"])
    assert rc == 0
    good = verify_item(bid, proof, comm, tc)
    assert oracle_verify(good) == 0 and be.blindbid_verify(good) == 0
    muts = mutations(bid, proof, comm, tc)
    items = [it for _, it in muts]
    gpu = be.blindbid_verify_each([good] + items)
    assert gpu[0] == 0
    for (name, it), g in zip(muts, gpu[1:]):
        o = oracle_verify(it)
        assert o != 0, name
        assert g == o, f"{name}: gpu {g} oracle {o}"
        assert be.blindbid_verify(it) == o, name


def test_non_member_bid_does_not_verify(be):
    """the prover does not check the witness: a bid whose x is not in the list yields a proof that fails to verify
    (SURVEY.md §8b), identically on both sides"""
    bid = make_case(77, 4)
    pl = bytearray(bid["pub_list"]); t = bid["toggle"]
    pl[32 * t:32 * t + 32] = orc.random_scalars(999, 1)
    bid["pub_list"] = bytes(pl)
    rc, proof, comm, tc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"])
    st, gproof, gcomm, gtc = be.blindbid_prove(bid)
    assert rc == 0 and st == 0 and (gproof, gcomm, gtc) == (proof, comm, tc)
    item = verify_item(bid, gproof, gcomm, gtc)
    assert oracle_verify(item) != 0 and be.blindbid_verify(item) == oracle_verify(item)


def test_toggle_beyond_list_and_non_canonical_inputs(be):
    """the reference accepts any toggle (`i as u64 == toggle`): an index beyond the list gives all-zero toggle bits and a
    proof that does not verify — same bytes as the oracle's; serde-decoded scalars must be canonical (FormatError)"""
    bid = make_case(95, 4)
    bid["toggle"] = 9
    rc, proof, comm, tc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"])
    st, gproof, gcomm, gtc = be.blindbid_prove(bid)
    assert rc == 0 and st == 0 and (gproof, gcomm, gtc) == (proof, comm, tc)
    item = verify_item(bid, gproof, gcomm, gtc)
    assert oracle_verify(item) != 0 and be.blindbid_verify(item) == oracle_verify(item)
    good = make_case(96, 2)
    bad = dict(good); bad["k"] = le(L_ORDER + 5)
    assert be.blindbid_prove(bad)[0] == -2
    st, p, c, t = be.blindbid_prove(good)
    it = verify_item(good, p, c, t)
    assert be.blindbid_verify(it) == 0
    it["seed"] = le(from_le(good["seed"]) + L_ORDER)
    assert be.blindbid_verify(it) == -2


def test_generator_capacity_limits(be):
    """L = 202 is the largest list that fits gens(2048, 1); L = 203 -> InvalidGeneratorsLength on both sides"""
    bid = make_case(90, 203)
    rc, _, _, _ = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"])
    st, _, _, _ = be.blindbid_prove(bid)
    assert rc == -1 and st == -1
    bid = make_case(91, 202)
    rc, proof, comm, tc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"])
    st, gproof, gcomm, gtc = be.blindbid_prove(bid)
    assert rc == 0 and st == 0 and (gproof, gcomm, gtc) == (proof, comm, tc)
    assert be.blindbid_verify(verify_item(bid, gproof, gcomm, gtc)) == 0


@pytest.mark.parametrize("part", [None, 7, -1])
def test_batch_verify_equals_and_of_singles(be, part, monkeypatch):
    """BASELINE config 4 at test size: combined check verdict == AND of single verdicts, with 0 / 1 / k bad proofs.
    part = 7 shrinks the 1024-request part size, so the batch is cut into four parts verified concurrently on sibling
    contexts ("lanes"), one combination per part. part = -1 keeps the whole group on one stream (default: the
    variable-base MSM over the requests' own points runs on a second stream beside the scalar assembly)."""
    if part == -1:
        monkeypatch.setenv("BBP_VERIFY_STREAMS", "1")
        part = None
    if part:
        monkeypatch.setenv("BBP_PROVE_PART", str(part))
        monkeypatch.setenv("BBP_PROVE_LANES", "3")
    n = 24
    cases = [make_case(200 + i, 8) for i in range(n)]
    outs = be.blindbid_prove_batch(cases)
    items = [verify_item(b, p, c, t, i) for i, (b, (st, p, c, t)) in enumerate(zip(cases, outs))]
    assert all(st == 0 for st, _, _, _ in outs)
    ok, st = be.blindbid_verify_batch(items, seed32("batch0"))
    assert ok and st == [0] * n
    for bad in ([5], [0, 7, 23]):
        its = [dict(x) for x in items]
        for b in bad:
            p = bytearray(its[b]["proof"])
            p[-1] ^= 1 if b != 7 else 0
            if b == 7:
                p[1 + 32 * 8] ^= 1       # t_x
            its[b]["proof"] = bytes(p)
        ok, st = be.blindbid_verify_batch(its, seed32("batch1"))
        singles = be.blindbid_verify_each(its)
        assert not ok
        assert st == singles
        assert [i for i, s in enumerate(st) if s != 0] == bad
        for b in bad:
            assert oracle_verify(its[b]) == st[b]
    # a request with a malformed proof is reported and does not poison the others
    its = [dict(x) for x in items]
    its[3]["proof"] = its[3]["proof"][:-1]
    ok, st = be.blindbid_verify_batch(its, seed32("batch2"))
    assert not ok and st[3] == -2 and all(s == 0 for i, s in enumerate(st) if i != 3)
    # a request with a point that does not decompress gets weight zero on the device and its own status; the rest pass
    its = [dict(x) for x in items]
    p = bytearray(its[11]["proof"]); p[1] |= 1; its[11]["proof"] = bytes(p)
    assert oracle_verify(its[11]) != 0
    ok, st = be.blindbid_verify_batch(its, seed32("batch3"))
    assert not ok and st[11] == oracle_verify(its[11]) and all(s == 0 for i, s in enumerate(st) if i != 11)
    assert st == be.blindbid_verify_each(its)


@pytest.mark.parametrize("regroup", ["8", "5", "0"])
def test_failed_combination_is_narrowed_by_runs(be, regroup, monkeypatch):
    """After a failed combined check the batch is re-combined in runs of g requests (BBP_VERIFY_REGROUP, default 8; 0 = off)
    and only the failing runs are checked request by request: the verdicts must be those of the per-request path, with a
    ragged last run (37 = 4 x 8 + 5), culprits in the first and the last run, and a request whose point does not
    decompress (weight zero: its run still passes)."""
    monkeypatch.setenv("BBP_VERIFY_REGROUP", regroup)
    n = 37
    cases = [make_case(900 + i, 2) for i in range(n)]
    outs = be.blindbid_prove_batch(cases)
    assert all(o[0] == 0 for o in outs)
    items = [verify_item(b, p, c, t, i) for i, (b, (st, p, c, t)) in enumerate(zip(cases, outs))]
    ok, st = be.blindbid_verify_batch(items, seed32("runs0"))
    assert ok and st == [0] * n
    its = [dict(x) for x in items]
    for b, pos in ((3, -1), (36, 1 + 32 * 8)):
        p = bytearray(its[b]["proof"]); p[pos] ^= 1; its[b]["proof"] = bytes(p)
    p = bytearray(its[10]["proof"]); p[1] |= 1; its[10]["proof"] = bytes(p)      # A_I1 is no longer a point encoding
    ok, st = be.blindbid_verify_batch(its, seed32("runs1"))
    assert not ok
    assert [i for i, s in enumerate(st) if s != 0] == [3, 10, 36]
    assert st == be.blindbid_verify_each(its)
    for b in (3, 10, 36):
        assert oracle_verify(its[b]) == st[b]
    # every request bad: every run fails, everyone is checked alone
    its = [dict(x) for x in items]
    for b in range(n):
        p = bytearray(its[b]["proof"]); p[-1] ^= 2; its[b]["proof"] = bytes(p)
    ok, st = be.blindbid_verify_batch(its, seed32("runs2"))
    assert not ok and all(s != 0 for s in st) and st == be.blindbid_verify_each(its)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 100, 1024])
def test_batch_weights_are_the_documented_shake_tree(be, n):
    """The weights of a batch verification's random linear combination are derived on the device (rng_kernels.cuh:
    k_batch_weight_digests / k_batch_weights); hashlib restates the two-level SHAKE256 tree."""
    r = orc.random_scalars(4242 + n, n)
    seed = seed32("weights")
    lbl = lambda t: t.encode().ljust(32, b"\0")
    G = (n + 31) // 32
    hs = b""
    for g in range(G):
        cnt = min(32, n - 32 * g)
        hs += hashlib.shake_256(lbl("bbp batch digest v1") + g.to_bytes(8, "little") + cnt.to_bytes(8, "little") + r[1024 * g:1024 * g + 32 * cnt]).digest(32)
    root = hashlib.shake_256(lbl("bbp batch root v1") + seed + n.to_bytes(8, "little") + hs).digest(32)
    want = b"".join(le(from_le(hashlib.shake_256(lbl("bbp batch weight v1") + root + i.to_bytes(8, "little")).digest(64)) % L_ORDER) for i in range(n))
    assert be.test_batch_weights(r, seed) == want


def test_batch_verify_config4_full_size(be):
    """BASELINE config 4 at full size: 1024 proofs in ONE combined mega-check; with 1 and with 16 corrupted proofs the
    verdicts must single out exactly the corrupted requests (re-check pass), and equal the per-request verdicts.
    Size-independent properties only (the oracle needs ~50 ms per verification): single-request verification on the
    GPU is pinned against the oracle by the tests above."""
    n = 1024
    cases = [make_case(5000 + i, 8) for i in range(n)]
    outs = be.blindbid_prove_batch(cases)
    assert all(st == 0 and len(p) == 1121 for st, p, _, _ in outs)
    items = [verify_item(b, p, c, t, i) for i, (b, (st, p, c, t)) in enumerate(zip(cases, outs))]
    ok, st = be.blindbid_verify_batch(items, seed32("full"))
    assert ok and st == [0] * n
    # determinism: a second proving pass yields identical bytes
    again = be.blindbid_prove_batch(cases[:64])
    assert [o[1] for o in again] == [o[1] for o in outs[:64]]
    # spot-check three of them against the oracle
    for i in (0, 511, 1023):
        assert oracle_verify(items[i]) == 0
    for bad in ([777], list(range(5, 1024, 64))):
        its = [dict(x) for x in items]
        for k, b in enumerate(bad):
            p = bytearray(its[b]["proof"])
            pos = [-1, 1 + 32 * 8, 40, -70, 1 + 32 * 9 + 3][k % 5]      # b, t_x, A_O1, a, t_x_blinding
            p[pos] ^= 4
            its[b]["proof"] = bytes(p)
        ok, st = be.blindbid_verify_batch(its, seed32("full-bad"))
        assert not ok
        assert [i for i, s in enumerate(st) if s != 0] == bad
        assert st == be.blindbid_verify_each(its)


def test_legacy_proof_layout(be):
    """R1CSProof::to_bytes risk R1: the legacy 14-point layout without the phase byte (1216 B) is selectable
    and byte-identical to the oracle's legacy form; the two layouts do not verify under each other's parser"""
    bid = make_case(42, 4)
    try:
        be.set_proof_format(0)
        st, proof, comm, tc = be.blindbid_prove(bid)
        rc, oproof, ocomm, otc = orc.blindbid_prove(bid, bid["blindings"], bid["rng_seed"], versioned=0)
        assert st == 0 and rc == 0 and len(proof) == 1216
        assert (proof, comm, tc) == (oproof, ocomm, otc)
        item = verify_item(bid, proof, comm, tc)
        assert be.blindbid_verify(item) == 0
        assert orc.blindbid_verify(proof, comm, tc, bid["q"], bid["z_img"], bid["seed"], bid["pub_list"], item["rng_seed"], versioned=0) == 0
        be.set_proof_format(1)
        assert be.blindbid_verify(item) != 0        # 1216 bytes under the versioned parser: format / verification error
        st, vproof, _, _ = be.blindbid_prove(bid)
        assert len(vproof) == 1121 and vproof[1:1 + 96] == proof[:96] and vproof[1 + 96:] == proof[192:]
    finally:
        be.set_proof_format(1)


def test_ct_commit_path_gives_identical_bytes(be, monkeypatch):
    """BBP_CT_COMMIT=1: the secret-scalar MSMs (V / T commitments, A_I1, A_O1, S1 — constant-time multiscalar_mul upstream)
    take the uniform digit-table path whatever the batch size; same group elements, same proof bytes"""
    cases = [make_case(880 + i, 8) for i in range(40)]
    plain = be.blindbid_prove_batch(cases)
    monkeypatch.setenv("BBP_CT_COMMIT", "1")
    l0 = be.launch_count()
    ct = be.blindbid_prove_batch(cases)
    assert ct == plain
    assert be.blindbid_prove(cases[3]) == plain[3]
    monkeypatch.delenv("BBP_CT_COMMIT")
    for i in (0, 39):
        rc, oproof, ocomm, otc = orc.blindbid_prove(cases[i], cases[i]["blindings"], cases[i]["rng_seed"])
        assert rc == 0 and ct[i] == (0, oproof, ocomm, otc)
