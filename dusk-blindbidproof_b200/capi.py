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
        if self.ctx:
            lib().bbp_free(self.ctx)
            self.ctx = ctypes.c_void_p()

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

    def msm_vartime_ptr(self, scalars_host_ptr, points_ext_host_ptr, n):
        """bbp_msm_vartime on raw host pointers (e.g. pinned torch tensors)."""
        out = _out(32)
        _chk(lib().bbp_msm_vartime(self.ctx, ctypes.c_void_p(scalars_host_ptr), ctypes.c_void_p(points_ext_host_ptr), _sz(n), out), "bbp_msm_vartime")
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
        return dict(c=out[0], W=out[1], S=out[2], CH=out[3])

    def int_peak(self):
        v = ctypes.c_double()
        _chk(lib().bbp_int_peak(self.ctx, ctypes.byref(v)), "bbp_int_peak")
        return v.value

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
