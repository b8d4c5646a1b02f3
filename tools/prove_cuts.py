"""Batched proving of B bids under different part sizes / lane counts (development aid):
python tools/prove_cuts.py B "part:lanes,part:lanes,..." (part 0 = library default)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader  # noqa: E402

pkg = bbp_loader.load()
from bench import synth_bid  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cuts = [tuple(int(x) for x in c.split(":")) for c in (sys.argv[2] if len(sys.argv) > 2 else "0:3").split(",")]
be = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
pp = pkg.capi.PreparedProve([synth_bid(pkg.capi, i, 8) for i in range(B)])
for part, lanes in cuts:
    if part:
        os.environ["BBP_PROVE_PART"] = str(part)
    else:
        os.environ.pop("BBP_PROVE_PART", None)
    os.environ["BBP_PROVE_LANES"] = str(lanes)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        be.blindbid_prove_prepared(pp)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"B={B} part={part or 'default'} lanes={lanes}: best {1e3 * ts[0]:.2f} ms median {1e3 * ts[len(ts) // 2]:.2f} ms = {B / ts[0]:.0f}/s", flush=True)
