"""N > 1 host-side logic on CPU (gloo, world_size 2): the point-range partition and the all-gather of 128-byte partial
sums that sharded MSM / batch verification use. The per-rank partials come from the oracle here (there
is no GPU on this box); the GPU path calls the same sharding helpers with NCCL."""
import ctypes
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    import bbp_loader
    sh = bbp_loader.load().sharding
    for n in (0, 1, 7, 1024, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            parts = [sh.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(10, 2, 2)


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bbp_loader
    import orc
    sh = bbp_loader.load().sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts = orc.random_points(7, n)
    scs = orc.random_scalars(7, n)
    a, b = sh.shard_range(n, rank, world)
    part_c = orc.msm(scs[32 * a:32 * b], pts[32 * a:32 * b], algo=1)
    ext = ctypes.create_string_buffer(128)
    assert orc.lib().orc_ge_decompress_ext(ext, part_c) == 1
    partial = torch.frombuffer(bytearray(ext.raw), dtype=torch.uint8)
    gathered = sh.gather_partials(dist, partial)
    assert gathered.numel() == 128 * world
    # every rank sums all partials locally and must get the full MSM
    comp = ctypes.create_string_buffer(32)
    parts = []
    for r in range(world):
        orc.lib().orc_ge_compress_ext(comp, bytes(gathered[128 * r:128 * (r + 1)].numpy()))
        parts.append(comp.raw)
    total = orc.msm(b"".join((1).to_bytes(32, "little") for _ in parts), b"".join(parts), algo=0)
    full = orc.msm(scs, pts, algo=1)
    own = bytes(gathered[128 * rank:128 * (rank + 1)].numpy()) == ext.raw
    q.put((rank, total == full, own))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_partials_world2_gloo():
    world, n = 2, 301
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)


class _FakeBackend:
    """stands in for the GPU backend on CPU tensors: the partial sums come from the oracle, everything else (row packing,
    the single all-gather, flag handling, the verdict) is the product's sharding.sharded_batch_verify"""

    def __init__(self, orc, partial_points, local_ok):
        self.orc, self.partial, self.local_ok = orc, partial_points, local_ok

    def blindbid_verify_batch_partial(self, items, seed, ptr):
        ctypes.memmove(ptr, self.partial, 256)
        return self.local_ok, []

    def sharded_verdict_device(self, rows_ptr, world, row_stride, out_ptr):
        # what k_sharded_verdict computes, with the oracle's group arithmetic: every flag set and the points sum to the identity
        comp = ctypes.create_string_buffer(32)
        cs, flags = [], []
        for r in range(world):
            row = ctypes.string_at(rows_ptr + r * row_stride, row_stride)
            flags.append(row[256])
            for k in range(2):
                self.orc.lib().orc_ge_compress_ext(comp, row[128 * k:128 * k + 128])
                cs.append(comp.raw)
        total = self.orc.msm(b"".join((1).to_bytes(32, "little") for _ in cs), b"".join(cs), algo=0)
        ctypes.memmove(out_ptr, bytes([1 if all(flags) and total == bytes(32) else 0]), 1)


def _verify_worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bbp_loader
    import orc
    sh = bbp_loader.load().sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 contributes (P, Q), rank 1 contributes (-P, -Q) [case "identity"] or (-P, Q) [case "nonzero"]
    P, Q = orc.random_points(3, 2)[:32], orc.random_points(3, 2)[32:]
    neg = (orc.L_ORDER - 1).to_bytes(32, "little")
    ext = ctypes.create_string_buffer(128)

    def ext_of(scalar, point):
        c = orc.msm(scalar, point, algo=0)
        assert orc.lib().orc_ge_decompress_ext(ext, c) == 1
        return ext.raw

    one = (1).to_bytes(32, "little")
    if rank == 0:
        partial = ext_of(one, P) + ext_of(one, Q)
    else:
        partial = ext_of(neg, P) + ext_of(one if case == "nonzero" else neg, Q)
    local_ok = not (case == "flag" and rank == 1)
    be = _FakeBackend(orc, partial, local_ok)
    d_out = torch.zeros(32, dtype=torch.uint8)
    verdict = sh.sharded_batch_verify(be, dist, [], bytes(32), None, None, d_out)
    q.put((rank, verdict))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,want", [("identity", True), ("nonzero", False), ("flag", False)])
def test_sharded_batch_verify_world2_gloo(case, want):
    """the sharded verdict over gloo, world size 2: partial sums that cancel -> accept; that do not -> reject; that cancel
    while one rank refused a request locally -> reject (the flag rides in the same all-gather row)"""
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_verify_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [v for _, v in sorted(res)] == [want, want]
