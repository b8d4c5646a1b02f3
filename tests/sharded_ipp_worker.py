"""torchrun worker of tests/test_gpu_rangeproof.py::test_sharded_ipp_two_ranks_nccl and of bench.py's sharded config-5 leg:
every rank proves the same aggregated range proofs cooperatively and compares with its own unsharded proof."""
import datetime
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bbp_loader  # noqa: E402

L_ORDER = 2**252 + 27742317777372353535851937790883648493


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    pkg = bbp_loader.load()
    be = pkg.Backend(device=local, gens_capacity=64, party_capacity=64)
    n = 2
    seeds = b"".join(hashlib.sha256(b"w%d" % i).digest() for i in range(n))
    vals = [[int.from_bytes(hashlib.shake_256(b"w-v" + bytes([i, k])).digest(8), "little") for i in range(64)] for k in range(n)]
    bls = b"".join((int.from_bytes(hashlib.shake_256(b"w-bl" + bytes([k, i])).digest(64), "little") % L_ORDER).to_bytes(32, "little")
                   for k in range(n) for i in range(64))
    st, plain, Vs = be.rangeproof_prove_batch(vals, bls, 64, 64, seeds)
    assert st == [0] * n
    stats = pkg.sharding.enable_sharded_ipp(be, dist)
    with torch.cuda.stream(torch.cuda.ExternalStream(be.stream())):
        st, sharded, Vs2 = be.rangeproof_prove_batch(vals, bls, 64, 64, seeds)
    pkg.sharding.disable_sharded_ipp(be)
    assert st == [0] * n and sharded == plain and Vs2 == Vs, "sharded proof differs from the single-GPU proof"
    assert stats["allgathers"] == 12 and stats["bytes_per_rank"] == 2 * n * 128
    assert be.rangeproof_verify_batch(sharded, Vs, 64, 64, bytes(32) * n) == [0] * n
    dist.barrier()
    print("SHARDED-IPP-OK rank", dist.get_rank(), flush=True)
    dist.destroy_process_group()


main()
