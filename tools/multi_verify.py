"""Per-rank batch verification / proving rate with N independent processes on one box, no collectives (development aid for host
contention): torchrun --nproc-per-node N tools/multi_verify.py [per_call=3072] [seconds=2]"""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bbp_loader
from bench import synth_bid

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
pkg = bbp_loader.load()
per_call = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
be = pkg.Backend(device=local, gens_capacity=2048, party_capacity=1)
bids = [synth_bid(pkg.capi, rank * 100000 + i, 8) for i in range(1024)]
outs = be.blindbid_prove_batch(bids)
items = [dict(proof=o[1], commitments=o[2], t_c=o[3], score=b["q"], z_img=b["z_img"], seed=b["seed"], pub_list=b["pub_list"],
              rng_seed=hashlib.sha256(b"v%d" % i).digest()) for i, (b, o) in enumerate(zip(bids, outs))]
items = (items * ((per_call + 1023) // 1024))[:per_call]
prep = pkg.capi.PreparedVerify(items)
for _ in range(3):
    ok, _ = be.blindbid_verify_batch(prep, bytes(32))
    assert ok
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
n = 0
while time.perf_counter() - t0 < secs:
    be.blindbid_verify_batch(prep, bytes(32))
    n += 1
dt = time.perf_counter() - t0
rate = n * per_call / dt
t = torch.tensor([rate], dtype=torch.float64)
if world > 1:
    dist.all_reduce(t)
if rank == 0:
    print(f"BBP_BLOCKING_SYNC={os.environ.get('BBP_BLOCKING_SYNC')} world {world} per_call {per_call}: {t.item():.0f} proofs/s total, rank 0 {1e3 * dt / n:.2f} ms per call", flush=True)
be.close()
