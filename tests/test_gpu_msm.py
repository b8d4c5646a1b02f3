"""GPU parity, layer K3: the Pippenger MSM behind the dalek trait surface (BASELINE config 2)
against the libsodium golden sums, the CPU oracle, and size-independent properties at the full 2^20 size.
Bit-exact: the output is a canonical compressed Ristretto point."""
import hashlib
import json
import os

import pytest

import orc
from orc import L_ORDER, from_le, le

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ristretto_libsodium.json")))
ZERO32 = bytes(32)


@pytest.fixture(scope="module")
def be():
    from gpu_util import backend
    return backend()


def test_golden_msm_libsodium(be):
    """Sums computed with libsodium scalarmult + add (tests/golden/make_golden.py)."""
    for g in GOLD["msm"]:
        n = g["n"]
        st = hashlib.shake_256(b"golden-msm" + n.to_bytes(4, "little")).digest(128 * n)
        pts = be.from_uniform_bytes(b"".join(st[128 * i:128 * i + 64] for i in range(n)))
        scs = b"".join(le(from_le(st[128 * i + 64:128 * i + 128]) % L_ORDER) for i in range(n))
        assert be.msm_optional(scs, pts).hex() == g["result"], n


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 32, 33, 189, 190, 191, 499, 500, 799, 800, 1023, 1024, 2933, 4143, 8286])
def test_msm_matches_oracle_ragged(be, n):
    """Ragged sizes around dalek's Straus/Pippenger thresholds (190, 500, 800) and the protocol's sizes."""
    from gpu_util import gpu_random_points
    pts = gpu_random_points(n, n)
    scs = orc.random_scalars(n, n)
    want = orc.msm(scs, pts, algo=1, threads=4)
    assert be.msm_optional(scs, pts) == want
    if n <= 33:
        assert orc.msm(scs, pts, algo=0) == want


@pytest.mark.parametrize("logn", [10, 12, 14, 16])
def test_msm_sweep_matches_oracle(be, logn):
    from gpu_util import gpu_random_points
    n = 1 << logn
    for seed in (0, 1):
        pts = gpu_random_points(seed, n)
        scs = orc.random_scalars(seed, n)
        tab, ok = be.points_from_compressed(pts)
        assert ok
        assert be.msm_points(scs, tab) == orc.msm(scs, pts, algo=1, threads=8)
        tab.free()


def test_msm_scalar_edge_cases(be):
    """zero scalars, l-1, unreduced 256-bit scalars (reduced mod l as dalek arithmetic does), identity points."""
    from gpu_util import gpu_random_points
    n = 300
    pts = bytearray(gpu_random_points(77, n))
    pts[32 * 5:32 * 6] = ZERO32            # identity encodings are valid points
    pts[32 * 6:32 * 7] = ZERO32
    pts = bytes(pts)
    scs = bytearray(orc.random_scalars(77, n))
    special = [0, 1, L_ORDER - 1, L_ORDER, L_ORDER + 1, 2**252, 2**255 - 1, 2**256 - 1, 2**253, 8 * L_ORDER + 3]
    for i, s in enumerate(special):
        scs[32 * (10 + i):32 * (11 + i)] = le(s)
    scs = bytes(scs)
    assert be.msm_optional(scs, pts) == orc.msm(scs, pts, algo=1)
    # all-zero scalars -> identity
    assert be.msm_optional(bytes(32 * n), pts) == ZERO32
    # one point, scalar 1 -> the point itself; scalar l -> identity
    one = pts[:32]
    assert be.msm_optional(le(1), one) == one
    assert be.msm_optional(le(L_ORDER), one) == ZERO32
    # P - P
    assert be.msm_optional(le(1) + le(L_ORDER - 1), one + one) == ZERO32


def test_msm_empty_and_bad_arguments(be):
    """the empty sum is the identity (dalek semantics); malformed calls return BBP_ERR_INPUT instead of crashing"""
    import ctypes
    import bbp_loader
    capi = bbp_loader.load().capi
    L = capi.lib()
    out = ctypes.create_string_buffer(b"\xff" * 32, 32)
    assert L.bbp_msm_optional(be.ctx, None, None, ctypes.c_size_t(0), out) == 0 and out.raw == ZERO32
    out = ctypes.create_string_buffer(b"\xff" * 32, 32)
    assert L.bbp_msm_vartime(be.ctx, None, None, ctypes.c_size_t(0), out) == 0 and out.raw == ZERO32
    assert L.bbp_msm_vartime(be.ctx, None, None, ctypes.c_size_t(3), out) == capi.BBP_ERR_INPUT
    assert L.bbp_msm_optional(None, b"", b"", ctypes.c_size_t(1), out) == capi.BBP_ERR_INPUT
    assert L.bbp_decompress(be.ctx, None, ctypes.c_size_t(1), out, None) == capi.BBP_ERR_INPUT
    h = ctypes.c_void_p()
    assert L.bbp_points_from_compressed(be.ctx, b"", ctypes.c_size_t(0), ctypes.byref(h), None) == capi.BBP_ERR_INPUT
    # scalars / points length mismatch with a resident table
    from gpu_util import gpu_random_points
    tab, ok = be.points_from_compressed(gpu_random_points(1, 8))
    assert L.bbp_msm_points(be.ctx, orc.random_scalars(1, 7), ctypes.c_size_t(7), tab.handle, out) == capi.BBP_ERR_INPUT
    tab.free()


def test_msm_optional_none_on_invalid_point(be):
    """optional_multiscalar_mul returns None when any point fails to decompress."""
    from gpu_util import gpu_random_points
    n = 64
    pts = bytearray(gpu_random_points(5, n))
    scs = orc.random_scalars(5, n)
    assert be.msm_optional(scs, bytes(pts)) is not None
    pts[32 * 17] |= 1                       # negative s: invalid encoding
    assert be.msm_optional(scs, bytes(pts)) is None


def test_msm_adversarial_bucket_skew(be):
    """all scalars equal / all points equal (SURVEY.md §8d config 2 adversarial set)."""
    from gpu_util import gpu_random_points
    n = 4096
    pts = gpu_random_points(9, n)
    s = orc.random_scalars(9, 1)
    assert be.msm_optional(s * n, pts) == orc.msm(s * n, pts, algo=1, threads=4)
    same = pts[:32] * n
    scs = orc.random_scalars(10, n)
    tot = sum(from_le(scs[32 * i:32 * i + 32]) for i in range(n)) % L_ORDER
    assert be.msm_optional(scs, same) == orc.msm(le(tot), pts[:32], algo=0)


def test_msm_batched_slots(be):
    """n_slots independent MSMs over the same resident bases in one launch == the single-slot results."""
    from gpu_util import gpu_random_points
    n, slots = 777, 5
    pts = gpu_random_points(21, n)
    tab, ok = be.points_from_compressed(pts)
    scs = orc.random_scalars(21, n * slots)
    got = be.msm_points_batched(scs, tab, slots)
    for k in range(slots):
        assert got[32 * k:32 * k + 32] == orc.msm(scs[32 * n * k:32 * n * (k + 1)], pts, algo=1, threads=4), k
    tab.free()


def test_msm_full_size_properties(be):
    """2^20 points (BASELINE config 2 upper end), checked through size-independent properties:
    (1) additivity over a partition: MSM(all) == sum of the 16 chunk MSMs (each chunk size is oracle-checked above);
    (2) linearity: MSM(a*s, P) == a * MSM(s, P);
    (3) folding: with bases repeating with period 4096, MSM == oracle MSM of the per-base scalar sums."""
    from gpu_util import gpu_random_points
    n, chunk = 1 << 20, 1 << 16
    pts = gpu_random_points(3, n)
    scs = orc.random_scalars(3, n)
    tab, ok = be.points_from_compressed(pts)
    assert ok
    whole = be.msm_points(scs, tab)
    tab.free()
    parts = [be.msm_optional(scs[32 * o:32 * (o + chunk)], pts[32 * o:32 * (o + chunk)]) for o in range(0, n, chunk)]
    assert orc.msm(b"".join(le(1) for _ in parts), b"".join(parts), algo=0) == whole
    # linearity
    a = from_le(orc.random_scalars(4, 1))
    scaled = b"".join(le(from_le(scs[32 * i:32 * i + 32]) * a % L_ORDER) for i in range(n))
    tab, _ = be.points_from_compressed(pts)
    assert be.msm_points(scaled, tab) == orc.msm(le(a), whole, algo=0)
    tab.free()
    # folding onto 4096 distinct bases
    period = 4096
    rep = pts[:32 * period] * (n // period)
    sums = [0] * period
    for i in range(n):
        sums[i % period] += from_le(scs[32 * i:32 * i + 32])
    folded = b"".join(le(s % L_ORDER) for s in sums)
    assert be.msm_optional(scs, rep) == orc.msm(folded, pts[:32 * period], algo=1, threads=8)


def test_msm_lanes_agree_and_overlap(be):
    """bbp_lane: sibling contexts of the same GPU give the same bytes as the owning context, sequentially and when
    three host threads call them at once (each lane has its own stream, engine and scratch); the owner's launch counter
    includes the lanes'."""
    import threading
    from gpu_util import gpu_random_points
    n = 5000
    pts = gpu_random_points(77, n)
    scs = orc.random_scalars(78, n)
    want = orc.msm(scs, pts)
    lanes = [be.lane(k) for k in range(3)]
    assert lanes[0] is be
    l0 = be.launch_count()
    assert [b.msm_optional(scs, pts) for b in lanes] == [want] * 3
    assert be.launch_count() - l0 >= 3 * 10
    got = [[None] * 4 for _ in lanes]

    def worker(k):
        for r in range(4):
            got[k][r] = lanes[k].msm_optional(scs, pts)

    th = [threading.Thread(target=worker, args=(k,)) for k in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert all(x == want for row in got for x in row)
    # a table built on the owner is usable from a lane once the owner has been synchronised
    ext, valid = be.decompress(pts)
    table = be.points_from_extended(ext)
    be.sync()
    assert lanes[2].msm_points(scs, table) == want
    table.free()
