"""Batch verification of 1024 proofs with 1 / 16 / 64 corrupted ones, for several run lengths of the narrowing pass
(BBP_VERIFY_REGROUP; 0 = per-request pass at once): python tools/verify_corrupted_time.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader
import bench

pkg = bbp_loader.load()
capi = pkg.capi
n, L = 1024, 8
be = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
bids = [bench.synth_bid(capi, 31000 + i, L) for i in range(n)]
outs = be.blindbid_prove_batch(bids)
items = [dict(proof=p[1], commitments=p[2], t_c=p[3], score=b["q"], z_img=b["z_img"], seed=b["seed"], pub_list=b["pub_list"], rng_seed=bytes([i & 255]) * 32)
         for i, (b, p) in enumerate(zip(bids, outs))]
seed = bytes(range(32))
for n_bad in (0, 1, 16, 64):
    its = [dict(x) for x in items]
    bad = sorted({(k * 61 + 7) % n for k in range(n_bad)})
    for k, b in enumerate(bad):
        pb = bytearray(its[b]["proof"]); pb[-1 - 32 * (k % 2)] ^= 1; its[b]["proof"] = bytes(pb)
    prep = capi.PreparedVerify(its)
    row = []
    for g in ("0", "4", "8", "16", "32"):
        os.environ["BBP_VERIFY_REGROUP"] = g
        ok, st = be.blindbid_verify_batch(prep, seed)
        assert [i for i, s in enumerate(st) if s] == bad and ok == (n_bad == 0)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            be.blindbid_verify_batch(prep, seed)
            best = min(best, time.perf_counter() - t0)
        row.append(f"g={g}: {best * 1e3:.2f} ms")
    print(f"{n_bad} bad: " + "  ".join(row), flush=True)
be.close()
