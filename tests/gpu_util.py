"""Shared helpers of the -m gpu parity tests: one backend per session, seeded inputs."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bbp_loader  # noqa: E402

_backend = None


def backend():
    """The product backend on cuda:0 with the blind-bid generator set resident (BulletproofGens::new(2048, 1))."""
    global _backend
    if _backend is None:
        pkg = bbp_loader.load()
        _backend = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
    return _backend


def shake(tag, n):
    return hashlib.shake_256(tag).digest(n)


def gpu_random_points(seed, n):
    """n uniform compressed points, hashed on the host and mapped to the group by the GPU codec (which
    test_gpu_primitives pins against the oracle and the libsodium golden vectors)."""
    stream = shake(b"bbp-bench-points" + seed.to_bytes(8, "little"), 64 * n)
    return backend().from_uniform_bytes(stream)


def pkg():
    return bbp_loader.load()
