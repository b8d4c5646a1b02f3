"""Warm per-phase trace of the batched prover / verifier (development aid): python tools/warm_trace.py L B [reps]
Runs every call `reps` times untraced, then once with BBP_TRACE=1 (the library reads the variable per call)."""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader  # noqa: E402

pkg = bbp_loader.load()
from bench import synth_bid  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
be = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
bids = [synth_bid(pkg.capi, i, L) for i in range(B)]


def timed(fn, name):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t0)
    print(f"{name} L={L} B={B}: best {1e3 * best:.2f} ms = {B / best:.0f}/s", flush=True)
    os.environ["BBP_TRACE"] = "1"
    fn()
    del os.environ["BBP_TRACE"]
    return r


pp = pkg.capi.PreparedProve(bids)
timed(lambda: be.blindbid_prove_prepared(pp), "prove")
outs = pp.results()
assert all(o[0] == 0 for o in outs)
items = [dict(proof=o[1], commitments=o[2], t_c=o[3], score=b["q"], z_img=b["z_img"], seed=b["seed"], pub_list=b["pub_list"],
              rng_seed=hashlib.sha256(b"v%d" % i).digest()) for i, (b, o) in enumerate(zip(bids, outs))]
pv = pkg.capi.PreparedVerify(items)
ok, _ = timed(lambda: be.blindbid_verify_batch(pv, bytes(32)), "verify_batch")
assert ok
