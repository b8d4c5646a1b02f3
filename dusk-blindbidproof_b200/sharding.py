"""Multi-GPU plumbing: one process per GPU, torch.distributed for the only exchange the path has.

Both sharded operations reduce to the same pattern: every rank leaves a few extended partial sums (128 B each) in its
own memory, one all-gather makes all of them visible everywhere, and each rank adds them up locally (point addition is
not an NCCL reduction). The traffic is 128-256 bytes per GPU, so the collective is latency-bound by construction.
"""
import torch


def shard_range(n, rank, world):
    """Contiguous, balanced partition of n units (points or proofs): returns (start, stop) of this rank's slice."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_partials(dist, partial, out=None):
    """all-gather of this rank's partial sums (uint8 tensor, k x 128 B) -> rank-major tensor of world * k x 128 B.
    `dist` is torch.distributed (or None for a single process)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return partial
    world = dist.get_world_size()
    if out is None:
        out = torch.empty(partial.numel() * world, dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous())
    return out


def sharded_msm(be, dist, d_scalars, n_local, table, d_ext, d_gather, d_out):
    """This rank's slice of a point-range sharded MSM: local Pippenger -> 128 B extended partial -> all-gather -> local
    sum + compression. All tensors are preallocated uint8 CUDA tensors; everything is enqueued on the backend's stream
    (the caller makes it torch's current stream)."""
    world = 1 if dist is None else dist.get_world_size()
    if world == 1:
        be.msm_points_device(d_scalars.data_ptr(), n_local, table, d_out.data_ptr(), None)
        return
    be.msm_points_device(d_scalars.data_ptr(), n_local, table, None, d_ext.data_ptr())
    gather_partials(dist, d_ext, d_gather)
    be.sum_compress_device(d_gather.data_ptr(), world, d_out.data_ptr())


_PART = 256            # two extended points (static | dynamic partial sums)
_ROW = _PART + 16      # + this rank's local flag, padded to a 16-byte multiple


def sharded_batch_verify(be, dist, items, batch_seed, d_partial, d_gather, d_out):
    """Proof-range sharded batch verification: items = this rank's requests. Returns True iff the combined mega-check
    over ALL ranks' requests is the identity AND every rank accepted all of its requests locally (every rank returns the
    same verdict). One collective and one device -> host read per call: the rank's local flag (requests refused on the
    host, or with a point that does not decompress, never enter a partial sum, so the identity test alone would accept
    them) travels in the same all-gather row as its 256-byte partial sums."""
    world = 1 if dist is None else dist.get_world_size()
    if world == 1:
        ok, _ = be.blindbid_verify_batch(items, batch_seed)
        return ok
    buf = getattr(be, "_shard_bufs", None)
    if buf is None or buf[0].device != d_out.device or buf[1].numel() != _ROW * world:
        dev = d_out.device
        pin = dev.type == "cuda"
        buf = (torch.zeros(_ROW, dtype=torch.uint8, device=dev), torch.zeros(_ROW * world, dtype=torch.uint8, device=dev),
               torch.zeros(16, dtype=torch.uint8, device=dev),
               torch.zeros(16, dtype=torch.uint8).pin_memory() if pin else torch.zeros(16, dtype=torch.uint8),
               torch.zeros(16, dtype=torch.uint8).pin_memory() if pin else torch.zeros(16, dtype=torch.uint8))
        be._shard_bufs = buf
    row, rows, res, h_res, h_flag = buf
    local_ok, _ = be.blindbid_verify_batch_partial(items, batch_seed, row.data_ptr())
    h_flag[0] = 1 if local_ok else 0
    row[_PART:].copy_(h_flag, non_blocking=True)
    dist.all_gather_into_tensor(rows, row)
    # one launch: sum of the 2 x world partial sums, identity test in extended coordinates, AND of the flags
    be.sharded_verdict_device(rows.data_ptr(), world, _ROW, res.data_ptr())
    h_res.copy_(res, non_blocking=True)
    if res.is_cuda:
        torch.cuda.current_stream().synchronize()
    return bool(h_res[0].item())


def combine_verdicts(dist, local_ok, d_out):
    """Verdict of a sharded batch verification: the summed partials compress to the identity (32 zero bytes in d_out) AND
    every rank's local flag is set. Requests a rank rejected before the combination (malformed proof, point that does not
    decompress) never enter its partial sum, so the identity test alone would accept them; the flag travels in one MIN
    all-reduce next to the all-gather of the partials."""
    flag = torch.tensor([1 if local_ok else 0], dtype=torch.int32, device=d_out.device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item()) and bytes(d_out.cpu().numpy()) == bytes(32)


# ---- sharded inner-product argument (BASELINE config 5 at N GPUs) ------------------------------------
class _DevicePtr:
    """a raw device pointer as a __cuda_array_interface__ object (zero-copy view for torch.as_tensor)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def enable_sharded_ipp(be, dist):
    """Every rank calls this once and then drives the SAME proving call with the SAME inputs: rank r computes the MSM terms of
    the generator columns i = r (mod world) and after each round's MSM the ranks exchange their partial sums (2 x 128 B per
    proof) with one all-gather, issued from the library through the callback built here. Returns the per-call statistics
    dict (number of all-gathers, bytes)."""
    import ctypes
    world, rank = dist.get_world_size(), dist.get_rank()
    stats = {"allgathers": 0, "bytes_per_rank": 0}
    views = {}
    be._ipp_views = views                         # dropped by disable_sharded_ipp, while the backend's stream still exists
    stream = torch.cuda.ExternalStream(be.stream())
    proto = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

    def allgather(user, send, recv, nbytes):
        try:
            key = (send, recv, nbytes)
            if key not in views:
                views[key] = (torch.as_tensor(_DevicePtr(send, nbytes), device="cuda"), torch.as_tensor(_DevicePtr(recv, nbytes * world), device="cuda"))
            s, r = views[key]
            with torch.cuda.stream(stream):      # ordered after the MSM kernels the library queued on its own stream
                dist.all_gather_into_tensor(r, s)
            stats["allgathers"] += 1
            stats["bytes_per_rank"] = nbytes
            return 0
        except Exception:                         # never let an exception cross the C frame
            import traceback
            traceback.print_exc()
            return 1

    cb = proto(allgather)
    be._ipp_allgather = cb                        # keeps the callback alive as long as the backend
    be.set_ipp_shard(rank, world, cb)
    return stats


def disable_sharded_ipp(be):
    """must run before the backend is closed: the zero-copy views were used on the backend's stream, and torch records an
    event on every stream a tensor was used on when it frees it"""
    be.set_ipp_shard(0, 1, None)
    torch.cuda.synchronize()
    views = getattr(be, "_ipp_views", None)
    if views is not None:
        views.clear()
