"""Repeated batch verification of one prepared request array (development aid): python tools/verify_trace.py [batch]"""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader
pkg = bbp_loader.load()
from bench import synth_bid
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
be = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
bids = [synth_bid(pkg.capi, i, 8) for i in range(B)]
outs = be.blindbid_prove_batch(bids)
items = [dict(proof=o[1], commitments=o[2], t_c=o[3], score=b["q"], z_img=b["z_img"], seed=b["seed"], pub_list=b["pub_list"],
              rng_seed=hashlib.sha256(b"v%d" % i).digest()) for i, (b, o) in enumerate(zip(bids, outs))]
prep = pkg.capi.PreparedVerify(items)
for k in range(4):
    t0 = time.perf_counter(); ok, _ = be.blindbid_verify_batch(prep, bytes(32)); print("ms", 1e3 * (time.perf_counter() - t0), ok, flush=True)
