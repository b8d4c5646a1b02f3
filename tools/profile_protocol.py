"""One batched prove or verify call between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...` launch lists
(development aid; numbers printed under a profiler are not bench values)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bbp_loader  # noqa: E402

pkg = bbp_loader.load()
from bench import synth_bid  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "prove"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
be = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
bids = [synth_bid(pkg.capi, i, 8) for i in range(B)]
outs = be.blindbid_prove_batch(bids)
assert all(o[0] == 0 for o in outs)
items = [dict(proof=o[1], commitments=o[2], t_c=o[3], score=b["q"], z_img=b["z_img"], seed=b["seed"], pub_list=b["pub_list"],
              rng_seed=hashlib.sha256(b"v%d" % i).digest()) for i, (b, o) in enumerate(zip(bids, outs))]
ok, _ = be.blindbid_verify_batch(items, bytes(32))
assert ok
rt = torch.cuda.cudart()
torch.cuda.synchronize()
rt.cudaProfilerStart()
if what == "prove":
    be.blindbid_prove_batch(bids)
else:
    ok, _ = be.blindbid_verify_batch(items, bytes(32))
    assert ok
torch.cuda.synchronize()
rt.cudaProfilerStop()
print(what, B, "done")
