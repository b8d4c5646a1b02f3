"""Sharded paths on ONE GPU: every "rank" is a separate context on cuda:0, the all-gather is a torch.cat. Checks the
multi-GPU algebra of SURVEY.md §8e (point-range sharded MSM, proof-range sharded batch verification) against the
single-context results; the collective itself is covered by tests/test_sharding_gloo.py and exercised with NCCL by
bench.py --gpus N."""
import hashlib

import pytest

import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ranks():
    import bbp_loader
    pkg = bbp_loader.load()
    ctxs = [pkg.Backend(device=0, gens_capacity=2048, party_capacity=1) for _ in range(3)]
    yield pkg, ctxs
    for c in ctxs:
        c.close()


def test_point_range_sharded_msm_equals_single(ranks):
    import torch
    pkg, ctxs = ranks
    from gpu_util import gpu_random_points
    n, world = 10000, 3
    pts = gpu_random_points(31, n)
    scs = orc.random_scalars(31, n)
    want = orc.msm(scs, pts, algo=1, threads=4)
    partials = []
    for r, be in enumerate(ctxs):
        a, b = pkg.sharding.shard_range(n, r, world)
        tab, ok = be.points_from_compressed(pts[32 * a:32 * b])
        assert ok
        d_sc = torch.frombuffer(bytearray(scs[32 * a:32 * b]), dtype=torch.uint8).cuda()
        d_ext = torch.zeros(128, dtype=torch.uint8, device="cuda")
        be.msm_points_device(d_sc.data_ptr(), b - a, tab, None, d_ext.data_ptr())
        be.sync()
        partials.append(d_ext.clone())
        tab.free()
    gathered = torch.cat(partials)
    d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
    ctxs[0].sum_compress_device(gathered.data_ptr(), world, d_out.data_ptr())
    ctxs[0].sync()
    assert bytes(d_out.cpu().numpy()) == want


@pytest.mark.parametrize("part", [None, 2])
def test_proof_range_sharded_batch_verify(ranks, part, monkeypatch):
    """part = 2 shrinks the 1024-request part size: every rank then verifies its 3 requests as two parts on two lanes
    and folds their partial sums into the one it contributes"""
    import torch
    if part:
        monkeypatch.setenv("BBP_PROVE_PART", str(part))
    pkg, ctxs = ranks
    world = 3
    L = 4
    bids = []
    for i in range(9):
        bid = orc.make_bid(700 + i, L, i % L)
        bid["blindings"] = orc.bid_blindings(700 + i, L)
        bid["rng_seed"] = hashlib.sha256(b"s%d" % i).digest()
        bids.append(bid)
    outs = ctxs[0].blindbid_prove_batch(bids)
    items = [dict(proof=p, commitments=c, t_c=t, score=b["q"], z_img=b["z_img"], seed=b["seed"], pub_list=b["pub_list"],
                  rng_seed=hashlib.sha256(b"v%d" % i).digest()) for i, (b, (st, p, c, t)) in enumerate(zip(bids, outs))]
    seed = hashlib.sha256(b"batch").digest()

    def run(its):
        partials, local = [], []
        for r, be in enumerate(ctxs):
            a, b = pkg.sharding.shard_range(len(its), r, world)
            d_part = torch.zeros(256, dtype=torch.uint8, device="cuda")
            ok, st = be.blindbid_verify_batch_partial(its[a:b], seed, d_part.data_ptr())
            be.sync()
            partials.append(d_part.clone())
            local.append(ok)
        d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
        ctxs[0].sum_compress_device(torch.cat(partials).data_ptr(), 2 * world, d_out.data_ptr())
        # the one-launch verdict the product path uses (rows = partial sums | flag byte | padding): identity test without
        # compression AND the flags; must agree with sum-and-compress + the host-side AND
        rows = torch.zeros(world, 272, dtype=torch.uint8, device="cuda")
        for r in range(world):
            rows[r, :256] = partials[r]
            rows[r, 256] = 1 if local[r] else 0
        d_v = torch.full((16,), 7, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctxs[0].sharded_verdict_device(rows.data_ptr(), world, 272, d_v.data_ptr())
        ctxs[0].sync()
        is_identity = bytes(d_out.cpu().numpy()) == bytes(32)
        assert int(d_v[0].item()) == (1 if is_identity and all(local) else 0)
        return is_identity, local

    ok, local = run(items)
    assert ok and all(local)
    bad = [dict(x) for x in items]
    p = bytearray(bad[4]["proof"]); p[-1] ^= 1; bad[4]["proof"] = bytes(p)
    ok, local = run(bad)
    assert not ok and local == [True, False, True]      # only the shard holding request 4 fails locally
    assert ctxs[1].blindbid_verify_each(bad)[4] != 0
    # requests rejected BEFORE the combination never enter a partial sum: the sum of the partials is still the identity,
    # so the verdict has to come from the per-rank flags (sharding.combine_verdicts ANDs them over the ranks)
    for mutate in ("malformed", "undecompressable"):
        bad = [dict(x) for x in items]
        p = bytearray(bad[7]["proof"])
        if mutate == "malformed":
            p = p[:-5]                                   # R1CSProof::from_bytes -> FormatError
        else:
            p[97:129] = b"\x01" + bytes(31)             # T_1 (after the phase byte and A_I1, A_O1, S1) := an odd s: no valid encoding
        bad[7]["proof"] = bytes(p)
        sum_is_identity, local = run(bad)
        assert local == [True, True, False], (mutate, local)
        assert not (sum_is_identity and all(local))
        want = ctxs[0].blindbid_verify_each(bad)
        assert want[7] != 0 and all(w == 0 for i, w in enumerate(want) if i != 7)
