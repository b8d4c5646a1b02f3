"""ctypes binding of libbbp_b200.so (include/bbp.h). No arithmetic lives here."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbbp_b200.so")

BBP_OK = 0
BBP_ERR_INVALID_GENERATORS_LENGTH = -1
BBP_ERR_FORMAT = -2
BBP_ERR_VERIFICATION = -3
BBP_ERR_INPUT = -10
BBP_ERR_DECOMPRESS = -11
BBP_ERR_CUDA = -100
BBP_ERR_NCCL = -101


class BbpError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"{what} failed with bbp_status {code}")
        self.code = code


_lib = None


def lib():
    """Loads the product library. There is no fallback: a missing build is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run build.sh (or __graft_entry__.build()) first; there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        L.bbp_launch_count.restype = ctypes.c_uint64
        L.bbp_stream.restype = ctypes.c_uint64
        L.bbp_points_len.restype = ctypes.c_size_t
        L.bbp_free.restype = None
        L.bbp_points_free.restype = None
        _lib = L
    return _lib


def _sz(n):
    return ctypes.c_size_t(n)


def _out(n):
    return ctypes.create_string_buffer(n)


def _chk(rc, what):
    if rc != 0:
        raise BbpError(rc, what)


class Points:
    """Device-resident base table (bbp_points)."""

    def __init__(self, backend, handle, n):
        self.backend, self.handle, self.n = backend, handle, n

    def free(self):
        if self.handle:
            lib().bbp_points_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Backend:
    """One bbp_ctx: one CUDA device, one stream, resident generator tables."""

    def __init__(self, device=0, gens_capacity=0, party_capacity=1):
        self.ctx = ctypes.c_void_p()
        rc = lib().bbp_init(ctypes.byref(self.ctx), int(device), ctypes.c_uint32(gens_capacity), ctypes.c_uint32(party_capacity))
        _chk(rc, "bbp_init")
        self.device = device

    def close(self):
        # device tensors cached on this object (sharding.py) were used on the context's stream: they must be released while
        # that stream still exists (torch records an event on every stream a tensor was used on when it frees it)
        for k in ("_shard_bufs", "_ipp_views", "_ipp_allgather"):
            self.__dict__.pop(k, None)
        if self.ctx:
            if getattr(self, "_owner", None) is None:   # lanes belong to the context that created them
                lib().bbp_free(self.ctx)
            self.ctx = ctypes.c_void_p()

    def lane(self, k):
        """k-th sibling context of the same GPU (bbp_lane): own stream, engine and scratch; k = 0 is this context."""
        if k == 0:
            return self
        h = ctypes.c_void_p()
        _chk(lib().bbp_lane(self.ctx, ctypes.c_uint32(k), ctypes.byref(h)), "bbp_lane")
        view = Backend.__new__(Backend)
        view.ctx, view.device, view._owner = h, self.device, self   # keeps the owner alive
        return view

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- bookkeeping
    def launch_count(self):
        return int(lib().bbp_launch_count(self.ctx))

    def stream(self):
        return int(lib().bbp_stream(self.ctx))

    def sync(self):
        _chk(lib().bbp_sync(self.ctx), "bbp_sync")

    # ---- generators
    def pedersen_gens(self):
        b, bb = _out(32), _out(32)
        _chk(lib().bbp_pedersen_gens(self.ctx, b, bb), "bbp_pedersen_gens")
        return b.raw, bb.raw

    def bulletproof_gens(self, which, party, first, count):
        out = _out(32 * count)
        _chk(lib().bbp_bulletproof_gens(self.ctx, ord(which), ctypes.c_uint32(party), ctypes.c_uint32(first), ctypes.c_uint32(count), out),
             "bbp_bulletproof_gens")
        return out.raw

    # ---- base tables
    def points_from_compressed(self, compressed):
        n = len(compressed) // 32
        h = ctypes.c_void_p()
        valid = ctypes.c_int(1)
        _chk(lib().bbp_points_from_compressed(self.ctx, compressed, _sz(n), ctypes.byref(h), ctypes.byref(valid)), "bbp_points_from_compressed")
        return Points(self, h, n), bool(valid.value)

    def points_from_extended(self, ext):
        n = len(ext) // 128
        h = ctypes.c_void_p()
        _chk(lib().bbp_points_from_extended(self.ctx, ext, _sz(n), ctypes.byref(h)), "bbp_points_from_extended")
        return Points(self, h, n)

    # ---- MSM
    def msm_points(self, scalars, points):
        out = _out(32)
        _chk(lib().bbp_msm_points(self.ctx, scalars, _sz(len(scalars) // 32), points.handle, out), "bbp_msm_points")
        return out.raw

    def msm_points_batched(self, scalars, points, n_slots):
        out = _out(32 * n_slots)
        _chk(lib().bbp_msm_points_batched(self.ctx, scalars, _sz(points.n), _sz(n_slots), points.handle, out), "bbp_msm_points_batched")
        return out.raw

    def msm_points_device(self, scalars_dev_ptr, n, points, out_dev_ptr=None, out_ext_dev_ptr=None):
        _chk(lib().bbp_msm_points_device(self.ctx, ctypes.c_void_p(scalars_dev_ptr), _sz(n), points.handle, ctypes.c_void_p(out_dev_ptr),
                                         ctypes.c_void_p(out_ext_dev_ptr)), "bbp_msm_points_device")

    def sum_compress_device(self, ext_dev_ptr, n, out_dev_ptr):
        _chk(lib().bbp_sum_compress_device(self.ctx, ctypes.c_void_p(ext_dev_ptr), _sz(n), ctypes.c_void_p(out_dev_ptr)), "bbp_sum_compress_device")

    def sharded_verdict_device(self, rows_dev_ptr, world, row_stride, out_dev_ptr):
        _chk(lib().bbp_sharded_verdict_device(self.ctx, ctypes.c_void_p(rows_dev_ptr), _sz(world), _sz(row_stride), ctypes.c_void_p(out_dev_ptr)),
             "bbp_sharded_verdict_device")

    def msm_vartime_ptr(self, scalars_host_ptr, points_ext_host_ptr, n):
        """bbp_msm_vartime on raw host pointers (e.g. pinned torch tensors)."""
        out = _out(32)
        _chk(lib().bbp_msm_vartime(self.ctx, ctypes.c_void_p(scalars_host_ptr), ctypes.c_void_p(points_ext_host_ptr), _sz(n), out), "bbp_msm_vartime")
        return out.raw

    def msm_optional_ptr(self, scalars_host_ptr, points_compressed_host_ptr, n):
        """bbp_msm_optional on raw host pointers; None when a point fails to decompress."""
        out = _out(32)
        rc = lib().bbp_msm_optional(self.ctx, ctypes.c_void_p(scalars_host_ptr), ctypes.c_void_p(points_compressed_host_ptr), _sz(n), out)
        if rc == BBP_ERR_DECOMPRESS:
            return None
        _chk(rc, "bbp_msm_optional")
        return out.raw

    # ---- measurement hooks
    def set_profiling(self, on):
        _chk(lib().bbp_set_profiling(self.ctx, int(on)), "bbp_set_profiling")

    def msm_stage_ms(self):
        ms = (ctypes.c_float * 7)()
        _chk(lib().bbp_msm_stage_ms(self.ctx, ms, _sz(7)), "bbp_msm_stage_ms")
        return list(ms)

    @staticmethod
    def msm_plan(n):
        out = (ctypes.c_uint32 * 4)()
        _chk(lib().bbp_msm_plan(_sz(n), out), "bbp_msm_plan")
        return dict(c=out[0], W=out[1], S=out[2], B=out[3])

    def int_peak(self):
        """(IMAD.WIDE per second sustained, IMAD.WIDE per SM clock per SM)"""
        v, c = ctypes.c_double(), ctypes.c_double()
        _chk(lib().bbp_int_peak(self.ctx, ctypes.byref(v), ctypes.byref(c)), "bbp_int_peak")
        return v.value, c.value

    def set_ipp_shard(self, rank, world, allgather=None, emulate=0):
        """bbp_set_ipp_shard: allgather = a ctypes callback (sharding.enable_sharded_ipp builds it over torch.distributed)"""
        _chk(lib().bbp_set_ipp_shard(self.ctx, ctypes.c_uint32(rank), ctypes.c_uint32(world), allgather, None, int(emulate)), "bbp_set_ipp_shard")

    def int_peak_pairs(self):
        """the same as mad.lo.cc / madc.hi pairs: (pairs per second sustained, pairs per SM clock per SM)"""
        v, c = ctypes.c_double(), ctypes.c_double()
        _chk(lib().bbp_int_peak_pairs(self.ctx, ctypes.byref(v), ctypes.byref(c)), "bbp_int_peak_pairs")
        return v.value, c.value

    def msm_vartime(self, scalars, points_ext):
        out = _out(32)
        _chk(lib().bbp_msm_vartime(self.ctx, scalars, points_ext, _sz(len(scalars) // 32), out), "bbp_msm_vartime")
        return out.raw

    def msm_optional(self, scalars, points_compressed):
        """Returns the compressed result, or None when a point fails to decompress (optional_multiscalar_mul -> None)."""
        out = _out(32)
        rc = lib().bbp_msm_optional(self.ctx, scalars, points_compressed, _sz(len(scalars) // 32), out)
        if rc == BBP_ERR_DECOMPRESS:
            return None
        _chk(rc, "bbp_msm_optional")
        return out.raw

    # ---- codecs
    def decompress(self, compressed):
        n = len(compressed) // 32
        ext, valid = _out(128 * n), _out(n)
        _chk(lib().bbp_decompress(self.ctx, compressed, _sz(n), ext, valid), "bbp_decompress")
        return ext.raw, valid.raw

    def compress(self, ext):
        n = len(ext) // 128
        out = _out(32 * n)
        _chk(lib().bbp_compress(self.ctx, ext, _sz(n), out), "bbp_compress")
        return out.raw

    def from_uniform_bytes(self, b64):
        n = len(b64) // 64
        out = _out(32 * n)
        _chk(lib().bbp_from_uniform_bytes(self.ctx, b64, _sz(n), out), "bbp_from_uniform_bytes")
        return out.raw

    # ---- unit-test hooks
    def test_fe(self, a, b, op):
        n = len(a) // 32
        out = _out(32 * n)
        _chk(lib().bbp_test_fe(self.ctx, a, b, _sz(n), int(op), out), "bbp_test_fe")
        return out.raw

    def test_ge(self, a, b, op):
        n = len(a) // 32
        out = _out(32 * n)
        _chk(lib().bbp_test_ge(self.ctx, a, b, _sz(n), int(op), out), "bbp_test_ge")
        return out.raw

    def test_batch_weights(self, r, batch_seed):
        n = len(r) // 32
        out = _out(32 * n)
        _chk(lib().bbp_test_batch_weights(self.ctx, r, _sz(n), batch_seed, out), "bbp_test_batch_weights")
        return out.raw


# ---------------------------------------------------------------------------------------------- generic bulletproofs surface
class CsStruct(ctypes.Structure):
    _fields_ = [("n_multipliers", ctypes.c_uint32), ("n_commitments", ctypes.c_uint32), ("n_constraints", ctypes.c_uint32),
                ("con_ptr", ctypes.POINTER(ctypes.c_uint32)), ("term_var", ctypes.POINTER(ctypes.c_uint32)), ("term_coeff", ctypes.c_char_p)]


class Transcript:
    """merlin::Transcript over the C ABI (bbp_transcript_*)."""

    def __init__(self, label=None, _handle=None):
        lib().bbp_transcript_new.restype = ctypes.c_void_p
        lib().bbp_transcript_clone.restype = ctypes.c_void_p
        self.h = _handle if _handle is not None else lib().bbp_transcript_new(label, _sz(len(label)))
        if not self.h:
            raise BbpError(BBP_ERR_INPUT, "bbp_transcript_new")

    def clone(self):
        return Transcript(_handle=lib().bbp_transcript_clone(ctypes.c_void_p(self.h)))

    def append_message(self, label, msg):
        _chk(lib().bbp_transcript_append_message(ctypes.c_void_p(self.h), label, _sz(len(label)), msg, _sz(len(msg))), "bbp_transcript_append_message")

    def append_u64(self, label, x):
        _chk(lib().bbp_transcript_append_u64(ctypes.c_void_p(self.h), label, _sz(len(label)), ctypes.c_uint64(x)), "bbp_transcript_append_u64")

    def challenge_bytes(self, label, n):
        out = _out(n)
        _chk(lib().bbp_transcript_challenge_bytes(ctypes.c_void_p(self.h), label, _sz(len(label)), out, _sz(n)), "bbp_transcript_challenge_bytes")
        return out.raw

    def __del__(self):
        try:
            lib().bbp_transcript_free.restype = None
            lib().bbp_transcript_free(ctypes.c_void_p(self.h))
        except Exception:
            pass


def _cs_struct(cs):
    """cs = dict(n_mul, m, con_ptr[q + 1], term_var[], term_coeff bytes) -> (CsStruct, keep-alive tuple)"""
    con_ptr = (ctypes.c_uint32 * len(cs["con_ptr"]))(*cs["con_ptr"])
    term_var = (ctypes.c_uint32 * max(1, len(cs["term_var"])))(*cs["term_var"])
    st = CsStruct(cs["n_mul"], cs["m"], len(cs["con_ptr"]) - 1, con_ptr, term_var, cs["term_coeff"])
    return st, (con_ptr, term_var)


def cs_shape(cs):
    """host-only bbp_cs_shape: (status, [multipliers, constraints, commitments, padded n, coefficient-table entries])"""
    st, keep = _cs_struct(cs)
    out = (ctypes.c_size_t * 5)()
    rc = lib().bbp_cs_shape(ctypes.byref(st), out)
    return rc, list(out)


def _generic_methods():
    def r1cs_prove(self, transcript, cs, a_L, a_R, a_O, v, v_blinding, rng_seed):
        """Prover::prove over a flattened circuit; returns (status, proof, V). The transcript is advanced."""
        st, keep = _cs_struct(cs)
        V = _out(32 * max(1, cs["m"]))
        proof = _out(8192)
        plen = ctypes.c_size_t(8192)
        rc = lib().bbp_r1cs_prove(self.ctx, ctypes.c_void_p(transcript.h), ctypes.byref(st), a_L, a_R, a_O, v, v_blinding, rng_seed, V, proof, ctypes.byref(plen))
        if rc != 0:
            return rc, None, None
        return 0, proof.raw[:plen.value], V.raw[:32 * cs["m"]]

    def r1cs_verify(self, transcript, cs, proof, V, rng_seed):
        st, keep = _cs_struct(cs)
        return lib().bbp_r1cs_verify(self.ctx, ctypes.c_void_p(transcript.h), ctypes.byref(st), proof, _sz(len(proof)), V, rng_seed)

    def ipp_create(self, transcript, w, Gf, Hf, a, b):
        n = len(a) // 32
        out = _out(64 * n.bit_length() + 64)
        olen = ctypes.c_size_t(len(out))
        _chk(lib().bbp_ipp_create(self.ctx, ctypes.c_void_p(transcript.h), w, Gf, Hf, a, b, _sz(n), out, ctypes.byref(olen)), "bbp_ipp_create")
        return out.raw[:olen.value]

    return dict(r1cs_prove=r1cs_prove, r1cs_verify=r1cs_verify, ipp_create=ipp_create)


# ---------------------------------------------------------------------------------------------- outer boundary: TLV codec
def wire_prove_request(scalars7, pub_list, toggle):
    """the request frame of opcode 1 (d, k, y, y_inv, q, z_img, seed as 7 x 32 bytes)"""
    out = _out(64 + 40 * (8 + len(pub_list) // 32))
    n = ctypes.c_size_t(len(out))
    _chk(lib().bbp_wire_encode_prove_request(scalars7, pub_list, _sz(len(pub_list) // 32), ctypes.c_uint64(toggle), out, ctypes.byref(n)), "bbp_wire_encode_prove_request")
    return out.raw[:n.value]


def wire_proof_blob(proof, commitments, t_c):
    out = _out(len(proof) + 40 * (4 + (len(commitments) + len(t_c)) // 32))
    n = ctypes.c_size_t(len(out))
    _chk(lib().bbp_wire_encode_proof_blob(proof, _sz(len(proof)), commitments, _sz(len(commitments) // 32), t_c, _sz(len(t_c) // 32), out, ctypes.byref(n)),
         "bbp_wire_encode_proof_blob")
    return out.raw[:n.value]


def wire_decode_proof_blob(blob):
    proof, comm, tc = _out(len(blob)), _out(len(blob)), _out(len(blob))
    pl, nc, nt = ctypes.c_size_t(len(blob)), ctypes.c_size_t(len(blob) // 32), ctypes.c_size_t(len(blob) // 32)
    _chk(lib().bbp_wire_decode_proof_blob(blob, _sz(len(blob)), proof, ctypes.byref(pl), comm, ctypes.byref(nc), tc, ctypes.byref(nt)), "bbp_wire_decode_proof_blob")
    return proof.raw[:pl.value], comm.raw[:32 * nc.value], tc.raw[:32 * nt.value]


def wire_verify_request(blob, score, z_img, seed, pub_list):
    out = _out(len(blob) + 200 + 40 * (len(pub_list) // 32))
    n = ctypes.c_size_t(len(out))
    _chk(lib().bbp_wire_encode_verify_request(blob, _sz(len(blob)), score, z_img, seed, pub_list, _sz(len(pub_list) // 32), out, ctypes.byref(n)),
         "bbp_wire_encode_verify_request")
    return out.raw[:n.value]


def wire_frame(buf):
    """(status, header length, payload length): status 1 complete, 0 incomplete, < 0 malformed"""
    h, p = ctypes.c_size_t(0), ctypes.c_size_t(0)
    st = lib().bbp_wire_frame_len(buf, _sz(len(buf)), ctypes.byref(h), ctypes.byref(p))
    return st, h.value, p.value


def wire_parse(payload):
    """(opcode or status, handle); the handle must go to wire_free"""
    h = ctypes.c_void_p()
    op = lib().bbp_wire_parse(payload, _sz(len(payload)), ctypes.byref(h))
    return op, h


def wire_free(h):
    lib().bbp_wire_request_free.restype = None
    lib().bbp_wire_request_free(h)


def wire_execute(backend, handles, seed32=None):
    """bbp_wire_execute over parsed requests; returns the list of reply frames (None = nothing is written)"""
    n = len(handles)
    arr = (ctypes.c_void_p * n)(*[h.value for h in handles])
    replies = (ctypes.c_void_p * n)()
    lens = (ctypes.c_size_t * n)()
    _chk(lib().bbp_wire_execute(backend.ctx, _sz(n), arr, seed32, replies, lens), "bbp_wire_execute")
    out = []
    lib().bbp_wire_reply_free.restype = None
    for i in range(n):
        if replies[i]:
            out.append(ctypes.string_at(replies[i], lens[i]))
            lib().bbp_wire_reply_free(ctypes.c_void_p(replies[i]))
        else:
            out.append(None)
    return out


# ---------------------------------------------------------------------------------------------- blind-bid entry points
class ProveReq(ctypes.Structure):
    _fields_ = [(n, ctypes.c_char_p) for n in ("d", "k", "y", "y_inv", "q", "z_img", "seed", "pub_list")] + [
        ("L", ctypes.c_size_t), ("toggle", ctypes.c_uint64), ("blindings", ctypes.c_char_p), ("rng_seed", ctypes.c_char_p),
        ("proof_out", ctypes.c_void_p), ("proof_cap", ctypes.c_size_t), ("proof_len", ctypes.c_size_t),
        ("commitments_out", ctypes.c_void_p), ("t_c_out", ctypes.c_void_p), ("status", ctypes.c_int)]


class VerifyReq(ctypes.Structure):
    _fields_ = [("proof", ctypes.c_char_p), ("proof_len", ctypes.c_size_t), ("commitments", ctypes.c_char_p), ("n_commitments", ctypes.c_size_t),
                ("t_c", ctypes.c_char_p), ("n_t_c", ctypes.c_size_t), ("score", ctypes.c_char_p), ("z_img", ctypes.c_char_p), ("seed", ctypes.c_char_p),
                ("pub_list", ctypes.c_char_p), ("L", ctypes.c_size_t), ("rng_seed", ctypes.c_char_p), ("status", ctypes.c_int)]


PROOF_CAP = 2048


def _prove_reqs(bids):
    """bids: dicts with d,k,y,y_inv,q,z_img,seed,pub_list (bytes), toggle (int), blindings (bytes), rng_seed (bytes)."""
    n = len(bids)
    arr = (ProveReq * n)()
    keep = []
    for i, b in enumerate(bids):
        L = len(b["pub_list"]) // 32
        proof, comm, tc = _out(PROOF_CAP), _out(4 * 32), _out(32 * max(L, 1))
        keep.append((proof, comm, tc))
        r = arr[i]
        for f in ("d", "k", "y", "y_inv", "q", "z_img", "seed", "pub_list", "blindings", "rng_seed"):
            setattr(r, f, b[f])
        r.L, r.toggle = L, b["toggle"]
        r.proof_out, r.proof_cap = ctypes.cast(proof, ctypes.c_void_p), PROOF_CAP
        r.commitments_out, r.t_c_out = ctypes.cast(comm, ctypes.c_void_p), ctypes.cast(tc, ctypes.c_void_p)
    return arr, keep


def _verify_reqs(items):
    """items: dicts with proof, commitments, t_c, score, z_img, seed, pub_list, rng_seed (bytes)."""
    n = len(items)
    arr = (VerifyReq * n)()
    for i, v in enumerate(items):
        r = arr[i]
        r.proof, r.proof_len = v["proof"], len(v["proof"])
        r.commitments, r.n_commitments = v["commitments"], len(v["commitments"]) // 32
        r.t_c, r.n_t_c = v["t_c"], len(v["t_c"]) // 32
        r.score, r.z_img, r.seed = v["score"], v["z_img"], v["seed"]
        r.pub_list, r.L = v["pub_list"], len(v["pub_list"]) // 32
        r.rng_seed = v["rng_seed"]
    return arr


class PreparedProve:
    """bbp_prove_req array + output buffers built once (the caller's host buffers); reusable across calls."""

    def __init__(self, bids):
        self.bids = bids
        self.arr, self.keep = _prove_reqs(bids)
        self.n = len(bids)

    def results(self):
        out = []
        for i, (proof, comm, tc) in enumerate(self.keep):
            L = len(self.bids[i]["pub_list"]) // 32
            out.append((self.arr[i].status, proof.raw[:self.arr[i].proof_len], comm.raw, tc.raw[:32 * L]))
        return out


class PreparedVerify:
    """bbp_verify_req array built once; reusable across calls."""

    def __init__(self, items):
        self.items = items            # keeps the byte strings alive
        self.arr = _verify_reqs(items)
        self.n = len(items)

    def statuses(self):
        return [self.arr[i].status for i in range(self.n)]


def _backend_methods():
    def set_proof_format(self, versioned):
        _chk(lib().bbp_set_proof_format(self.ctx, int(versioned)), "bbp_set_proof_format")

    def blindbid_prove_batch(self, bids):
        """Proof::prove for a list of bids (or a PreparedProve) in one GPU pass. Returns [(status, proof, commitments, t_c)]."""
        prep = bids if isinstance(bids, PreparedProve) else PreparedProve(bids)
        _chk(lib().bbp_blindbid_prove_batch(self.ctx, _sz(prep.n), prep.arr), "bbp_blindbid_prove_batch")
        return prep.results()

    def blindbid_prove_prepared(self, prep):
        """the bare C call on a PreparedProve (results stay in its buffers: prep.results())"""
        _chk(lib().bbp_blindbid_prove_batch(self.ctx, _sz(prep.n), prep.arr), "bbp_blindbid_prove_batch")

    def blindbid_prove(self, bid):
        return self.blindbid_prove_batch([bid])[0]

    def blindbid_verify_each(self, items):
        """Verify::verify for every item independently; returns the list of statuses (0 = accept)."""
        arr = _verify_reqs(items)
        _chk(lib().bbp_blindbid_verify_each(self.ctx, _sz(len(items)), arr), "bbp_blindbid_verify_each")
        return [arr[i].status for i in range(len(items))]

    def blindbid_verify(self, item):
        return lib().bbp_blindbid_verify(self.ctx, item["proof"], _sz(len(item["proof"])), item["commitments"], _sz(len(item["commitments"]) // 32),
                                         item["t_c"], _sz(len(item["t_c"]) // 32), item["score"], item["z_img"], item["seed"], item["pub_list"],
                                         _sz(len(item["pub_list"]) // 32), item["rng_seed"])

    def blindbid_verify_batch(self, items, batch_seed):
        """One combined mega-check over a list of items (or a PreparedVerify); returns (all_ok, statuses)."""
        prep = items if isinstance(items, PreparedVerify) else PreparedVerify(items)
        ok = ctypes.c_int(0)
        _chk(lib().bbp_blindbid_verify_batch(self.ctx, _sz(prep.n), prep.arr, batch_seed, ctypes.byref(ok)), "bbp_blindbid_verify_batch")
        return bool(ok.value), prep.statuses()

    def blindbid_verify_batch_partial(self, items, batch_seed, partial_dev_ptr):
        prep = items if isinstance(items, PreparedVerify) else PreparedVerify(items)
        ok = ctypes.c_int(0)
        _chk(lib().bbp_blindbid_verify_batch_partial(self.ctx, _sz(prep.n), prep.arr, batch_seed, ctypes.c_void_p(partial_dev_ptr), ctypes.byref(ok)),
             "bbp_blindbid_verify_batch_partial")
        return bool(ok.value), prep.statuses()

    def pedersen_commit(self, values, blindings):
        n = len(values) // 32
        out = _out(32 * n)
        _chk(lib().bbp_pedersen_commit(self.ctx, values, blindings, _sz(n), out), "bbp_pedersen_commit")
        return out.raw

    def msm_gens(self, scalars, slot_len, n_slots):
        out = _out(32 * n_slots)
        _chk(lib().bbp_msm_gens(self.ctx, scalars, _sz(slot_len), _sz(n_slots), out), "bbp_msm_gens")
        return out.raw

    for f in (set_proof_format, blindbid_prove_batch, blindbid_prove_prepared, blindbid_prove, blindbid_verify_each, blindbid_verify, blindbid_verify_batch,
              blindbid_verify_batch_partial, pedersen_commit, msm_gens):
        setattr(Backend, f.__name__, f)


_backend_methods()
for _name, _fn in _generic_methods().items():
    setattr(Backend, _name, _fn)


# host-only helpers of the library (no GPU needed)
def mimc_hash(left, right):
    out = _out(32)
    _chk(lib().bbp_mimc_hash(left, right, out), "bbp_mimc_hash")
    return out.raw


def mimc_constants():
    out = _out(90 * 32)
    _chk(lib().bbp_mimc_constants(out), "bbp_mimc_constants")
    return out.raw


def circuit_shape(n_commitments, n_toggles):
    out = (ctypes.c_size_t * 3)()
    _chk(lib().bbp_blindbid_circuit_shape(_sz(n_commitments), _sz(n_toggles), out), "bbp_blindbid_circuit_shape")
    return out[0], out[1], out[2]


def _rangeproof_methods():
    def rangeproof_prove_batch(self, values, blindings, m, nbits, rng_seeds):
        """values: list of n_proofs lists of m ints; blindings: n_proofs*m*32 bytes; rng_seeds: n_proofs*32 bytes.
        Returns (statuses, [proof bytes], [commitment bytes])."""
        n = len(values)
        flat = (ctypes.c_uint64 * (n * m))(*[v for row in values for v in row])
        stride = 32 * (9 + 2 * 16)
        proofs, plen, comm = _out(stride * n), ctypes.c_size_t(0), _out(32 * m * n)
        st = (ctypes.c_int * n)()
        _chk(lib().bbp_rangeproof_prove_batch(self.ctx, _sz(n), flat, blindings, _sz(m), _sz(nbits), rng_seeds, proofs, _sz(stride), ctypes.byref(plen), comm, st),
             "bbp_rangeproof_prove_batch")
        L = plen.value
        return list(st), [proofs.raw[i * stride:i * stride + L] for i in range(n)], [comm.raw[32 * m * i:32 * m * (i + 1)] for i in range(n)]

    def rangeproof_verify_batch(self, proofs, commitments, m, nbits, rng_seeds):
        n = len(proofs)
        L = len(proofs[0])
        st = (ctypes.c_int * n)()
        _chk(lib().bbp_rangeproof_verify_batch(self.ctx, _sz(n), b"".join(proofs), _sz(L), _sz(L), b"".join(commitments), _sz(m), _sz(nbits), rng_seeds, st),
             "bbp_rangeproof_verify_batch")
        return list(st)

    def rangeproof_prove(self, values, blindings, nbits, rng_seed):
        m = len(values)
        flat = (ctypes.c_uint64 * m)(*values)
        proof, plen, comm = _out(2048), ctypes.c_size_t(2048), _out(32 * m)
        rc = lib().bbp_rangeproof_prove_multiple(self.ctx, flat, blindings, _sz(m), _sz(nbits), rng_seed, proof, ctypes.byref(plen), comm)
        return rc, proof.raw[:plen.value] if rc == 0 else b"", comm.raw

    def rangeproof_verify(self, proof, commitments, nbits, rng_seed):
        return lib().bbp_rangeproof_verify_multiple(self.ctx, proof, _sz(len(proof)), commitments, _sz(len(commitments) // 32), _sz(nbits), rng_seed)

    for f in (rangeproof_prove_batch, rangeproof_verify_batch, rangeproof_prove, rangeproof_verify):
        setattr(Backend, f.__name__, f)


_rangeproof_methods()
