"""torchrun worker (development aid): where the time of a sharded batch verification goes at N > 1."""
import datetime
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bbp_loader  # noqa: E402
from bench import synth_bid  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = bbp_loader.load()
    be = pkg.Backend(device=local, gens_capacity=2048, party_capacity=1)
    B = 1024
    bids = [synth_bid(pkg.capi, rank * 100000 + i, 8) for i in range(B)]
    outs = be.blindbid_prove_batch(bids)
    items = pkg.capi.PreparedVerify([dict(proof=o[1], commitments=o[2], t_c=o[3], score=b["q"], z_img=b["z_img"], seed=b["seed"], pub_list=b["pub_list"],
                                          rng_seed=hashlib.sha256(b"v%d" % i).digest()) for i, (b, o) in enumerate(zip(bids, outs))])
    seed = bytes(32)
    stream = torch.cuda.ExternalStream(be.stream(), device=local)
    d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
    row = torch.zeros(272, dtype=torch.uint8, device="cuda")
    rows = torch.zeros(272 * world, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(stream):
        for _ in range(3):
            assert pkg.sharding.sharded_batch_verify(be, dist, items, seed, None, None, d_out)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            pkg.sharding.sharded_batch_verify(be, dist, items, seed, None, None, d_out)
        torch.cuda.synchronize()
        full = (time.perf_counter() - t0) / 10
        # phases
        acc = [0.0] * 4
        for _ in range(10):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            be.blindbid_verify_batch_partial(items, seed, row.data_ptr())
            t1 = time.perf_counter()
            dist.all_gather_into_tensor(rows, row)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            acc[0] += t1 - t0; acc[1] += t2 - t1
        # local (non-sharded) call for comparison
        t0 = time.perf_counter()
        for _ in range(10):
            be.blindbid_verify_batch(items, seed)
        local_s = (time.perf_counter() - t0) / 10
    print(f"rank {rank}/{world}: sharded call {1e3 * full:.2f} ms; partial {1e3 * acc[0] / 10:.2f} ms, all_gather+sync {1e3 * acc[1] / 10:.3f} ms; plain local call {1e3 * local_s:.2f} ms", flush=True)
    dist.barrier()
    be.__dict__.pop("_shard_bufs", None)
    dist.destroy_process_group()


main()
