import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader
pkg = bbp_loader.load()
from bench import synth_bid
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
be = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
bids = [synth_bid(pkg.capi, i, 8) for i in range(B)]
for k in range(3):
    t0 = time.perf_counter(); be.blindbid_prove_batch(bids); print("ms", 1e3 * (time.perf_counter() - t0), flush=True)
