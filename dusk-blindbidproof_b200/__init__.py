"""B200 backend for Dusk's blind-bid Bulletproofs hot path.

The product is the C-ABI shared library ``libbbp_b200.so`` (include/bbp.h, csrc/). This package is the thin ctypes
view of it that tests/, bench.py and __graft_entry__.py drive; it carries no arithmetic of its own and has no CPU
fallback: importing ``capi`` without the built library, or creating a context without a CUDA device, fails loudly.

The directory name contains a dash (it follows the reference's name), so import it with ``load()`` from
``bbp_loader.py`` at the repo root, which registers it as ``dusk_blindbidproof_b200``.
"""
from . import capi, sharding  # noqa: F401
from .capi import Backend, BbpError, LIB_PATH  # noqa: F401
