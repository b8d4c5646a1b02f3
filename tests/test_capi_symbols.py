"""CPU-side checks of the drop-in boundary: libbbp_b200.so loads, exports every symbol include/bbp.h declares, and
refuses to run without a GPU (no CPU fallback). No compute calls here."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            syms += re.findall(r"\b(bbp_[a-z0-9_]+)\s*\(", src)
    return sorted(set(syms))


def test_library_exports_every_declared_symbol():
    import bbp_loader
    pkg = bbp_loader.load()
    L = pkg.capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import bbp_loader
    pkg = bbp_loader.load()
    with pytest.raises(pkg.BbpError) as e:
        pkg.Backend(device=0)
    assert e.value.code == pkg.capi.BBP_ERR_CUDA


def test_product_does_not_link_or_reference_the_oracle():
    """The product library must not depend on oracle/ (parity claims are void otherwise)."""
    so = os.path.join(ROOT, "dusk-blindbidproof_b200", "libbbp_b200.so")
    out = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "liboracle" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dusk-blindbidproof_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".inc", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "oracle/" not in text and "import orc" not in text, os.path.join(dirpath, f)


def test_rust_sys_crate_matches_header():
    """rust/bbp-sys is source only (no Rust toolchain here): at least keep its extern block in lock-step with include/bbp.h"""
    src = open(os.path.join(ROOT, "rust", "bbp-sys", "src", "lib.rs")).read()
    rust = set(re.findall(r"pub fn (bbp_[a-z0-9_]+)\s*\(", src))
    hdr = set(declared_symbols())
    assert rust and rust <= hdr, sorted(rust - hdr)
    # every entry point of the blind-bid path is bound
    for must in ("bbp_init", "bbp_free", "bbp_blindbid_prove_batch", "bbp_blindbid_verify_each", "bbp_blindbid_verify_batch",
                 "bbp_msm_vartime", "bbp_msm_optional", "bbp_rangeproof_prove_multiple", "bbp_rangeproof_verify_multiple"):
        assert must in rust, must
