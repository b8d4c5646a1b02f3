"""Overlap host and GPU phases by proving sub-batches on several contexts (one thread each) of the same GPU."""
import sys, os, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bbp_loader
pkg = bbp_loader.load()
from bench import synth_bid

L = 8
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
bids = [synth_bid(pkg.capi, i, L) for i in range(N)]
for lanes in (1, 2, 3):
    ctxs = [pkg.Backend(device=0, gens_capacity=2048, party_capacity=1) for _ in range(lanes)]
    for c in ctxs:
        c.blindbid_prove_batch(bids[:N // lanes])    # warm (allocations)
    for chunk in (N // lanes,):
        res = [None] * lanes

        def work(k):
            mine = bids[k::lanes]
            out = []
            for off in range(0, len(mine), chunk):
                out += ctxs[k].blindbid_prove_batch(mine[off:off + chunk])
            res[k] = out

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(k,)) for k in range(lanes)]
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        assert all(o[0] == 0 for r in res for o in r)
        print(f"lanes={lanes} chunk={chunk}: {N} proofs in {dt*1e3:.1f} ms -> {N/dt:.0f} proofs/s", flush=True)
    for c in ctxs:
        c.close()
