"""ctypes view of the CPU oracle (oracle/liboracle.so). Test infrastructure only."""
import ctypes
import hashlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

L_ORDER = 2**252 + 27742317777372353535851937790883648493
P = 2**255 - 19

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.orc_msm_ext.restype = ctypes.c_double
    return _lib


def _buf(n):
    return ctypes.create_string_buffer(n)


def le(x, n=32):
    return int(x).to_bytes(n, "little")


def from_le(b):
    return int.from_bytes(b, "little")


def fe_op(name, *args):
    out = _buf(32)
    getattr(lib(), name)(out, *args)
    return out.raw


def sc_mul(a, b):
    return fe_op("orc_sc_mul", a, b)


def sc_from_wide(b64):
    return fe_op("orc_sc_from_wide", b64)


def shake(data, n):
    return hashlib.shake_256(data).digest(n)


def random_scalars(seed, n):
    """n uniform scalars: 64-byte SHAKE256 blocks wide-reduced mod l."""
    stream = shake(b"bbp-bench-scalars" + seed.to_bytes(8, "little"), 64 * n)
    return b"".join(le(from_le(stream[64 * i:64 * i + 64]) % L_ORDER) for i in range(n))


def random_points(seed, n):
    """n uniform group elements (compressed): from_uniform_bytes of SHAKE256 blocks."""
    stream = shake(b"bbp-bench-points" + seed.to_bytes(8, "little"), 64 * n)
    out = _buf(32)
    res = []
    for i in range(n):
        lib().orc_ge_from_uniform(out, stream[64 * i:64 * i + 64])
        res.append(out.raw)
    return b"".join(res)


def msm(scalars, points, algo=1, threads=1):
    n = len(scalars) // 32
    out = _buf(32)
    ok = lib().orc_msm(out, scalars, points, ctypes.c_size_t(n), algo, threads)
    return out.raw if ok else None


def mimc_hash(left, right):
    return fe_op("orc_mimc_hash", left, right)


def sc_invert(a):
    return fe_op("orc_sc_invert", a)


def make_bid(seed_int, L, toggle):
    """Synthetic bid: returns dict of 32-byte scalars + pub_list bytes."""
    st = shake(b"bbp-bid" + seed_int.to_bytes(8, "little"), 64 * (3 + L) + 8)
    k = le(from_le(st[0:64]) % L_ORDER)
    d = le(from_le(st[64:72]))
    seed = le(from_le(st[128:192]) % L_ORDER)
    zero = le(0)
    m = mimc_hash(k, zero)
    x = mimc_hash(d, m)
    y = mimc_hash(seed, x)
    z_img = mimc_hash(seed, m)
    y_inv = sc_invert(y)
    q = sc_mul(d, y_inv)
    pub = [le(from_le(st[64 * (3 + i):64 * (4 + i)]) % L_ORDER) for i in range(L)]
    pub[toggle] = x
    return dict(d=d, k=k, y=y, y_inv=y_inv, q=q, z_img=z_img, seed=seed, pub_list=b"".join(pub), L=L, toggle=toggle, x=x)


def bid_blindings(seed_int, L):
    st = shake(b"bbp-blindings" + seed_int.to_bytes(8, "little"), 64 * (4 + L))
    return b"".join(le(from_le(st[64 * i:64 * i + 64]) % L_ORDER) for i in range(4 + L))


def blindbid_prove(bid, blindings, rng32, versioned=1):
    L = bid["L"]
    proof = _buf(2048)
    plen = ctypes.c_size_t(2048)
    comm = _buf(4 * 32)
    tc = _buf(32 * L)
    rc = lib().orc_blindbid_prove(bid["d"], bid["k"], bid["y"], bid["y_inv"], bid["q"], bid["z_img"], bid["seed"], bid["pub_list"],
                                  ctypes.c_size_t(L), ctypes.c_uint64(bid["toggle"]), blindings, rng32, versioned, proof,
                                  ctypes.byref(plen), comm, tc)
    if rc != 0:
        return rc, None, None, None
    return 0, proof.raw[:plen.value], comm.raw, tc.raw


def blindbid_verify(proof, comm, tc, score, z_img, seed, pub_list, rng32, versioned=1, threads=1, want_mega=False):
    L = len(pub_list) // 32
    mega = _buf(32 * 8192) if want_mega else None
    nm = ctypes.c_size_t(8192)
    rc = lib().orc_blindbid_verify(proof, ctypes.c_size_t(len(proof)), versioned, comm, ctypes.c_size_t(len(comm) // 32), tc,
                                   ctypes.c_size_t(len(tc) // 32), score, z_img, seed, pub_list, ctypes.c_size_t(L), rng32, threads,
                                   mega, ctypes.byref(nm))
    if want_mega:
        return rc, mega.raw[:32 * nm.value]
    return rc


# ---- generic R1CS over flattened circuits, standalone inner-product argument (oracle side of tests/test_gpu_r1cs.py)
def _flat_args(cs):
    import array
    con_ptr = (ctypes.c_uint32 * len(cs["con_ptr"]))(*cs["con_ptr"])
    term_var = (ctypes.c_uint32 * max(1, len(cs["term_var"])))(*cs["term_var"])
    return con_ptr, term_var


def r1cs_prove_flat(label, gens_capacity, cs, a_L, a_R, a_O, v, v_blinding, rng32, versioned=1):
    """cs = dict(n_mul, m, con_ptr[q+1], term_var[], term_coeff bytes). Returns (rc, proof, V, transcript challenge after)."""
    con_ptr, term_var = _flat_args(cs)
    m, q = cs["m"], len(cs["con_ptr"]) - 1
    V = _buf(32 * max(1, m))
    proof = _buf(4096)
    plen = ctypes.c_size_t(4096)
    after = _buf(32)
    rc = lib().orc_r1cs_prove_flat(label, ctypes.c_size_t(len(label)), ctypes.c_size_t(gens_capacity), ctypes.c_size_t(cs["n_mul"]), ctypes.c_size_t(m),
                                   ctypes.c_size_t(q), con_ptr, term_var, cs["term_coeff"], a_L, a_R, a_O, v, v_blinding, rng32, versioned, V, proof,
                                   ctypes.byref(plen), after)
    if rc != 0:
        return rc, None, None, None
    return 0, proof.raw[:plen.value], V.raw[:32 * m], after.raw


def r1cs_verify_flat(label, gens_capacity, cs, proof, V, rng32, versioned=1):
    con_ptr, term_var = _flat_args(cs)
    m, q = cs["m"], len(cs["con_ptr"]) - 1
    after = _buf(32)
    rc = lib().orc_r1cs_verify_flat(label, ctypes.c_size_t(len(label)), ctypes.c_size_t(gens_capacity), ctypes.c_size_t(cs["n_mul"]), ctypes.c_size_t(m),
                                    ctypes.c_size_t(q), con_ptr, term_var, cs["term_coeff"], proof, ctypes.c_size_t(len(proof)), versioned, V, rng32, after)
    return rc, after.raw


def ipp_create(label, w, Gf, Hf, a, b):
    n = len(a) // 32
    lg = n.bit_length() - 1
    out = _buf(64 * lg + 64)
    after = _buf(32)
    lib().orc_ipp_create(label, ctypes.c_size_t(len(label)), w, Gf, Hf, a, b, ctypes.c_size_t(n), out, after)
    return out.raw, after.raw
