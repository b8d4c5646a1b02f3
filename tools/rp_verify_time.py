"""Aggregated range-proof verification (m = 64, n = 64), host vs device transcript replay: python tools/rp_verify_time.py [n_proofs]"""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader

pkg = bbp_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = nbits = 64
be = pkg.Backend(device=0, gens_capacity=nbits, party_capacity=m)
vals = [[int.from_bytes(hashlib.shake_256(b"v%d.%d" % (k, i)).digest(8), "little") for i in range(m)] for k in range(n)]
order = 2**252 + 27742317777372353535851937790883648493
bls = b"".join((int.from_bytes(hashlib.shake_256(b"b%d.%d" % (k, i)).digest(64), "little") % order).to_bytes(32, "little") for k in range(n) for i in range(m))
seeds = b"".join(hashlib.sha256(b"s%d" % k).digest() for k in range(n))
st, proofs, Vs = be.rangeproof_prove_batch(vals, bls, m, nbits, seeds)
assert st == [0] * n
vseeds = bytes(range(32)) * n
bad = list(proofs)
bad[n // 2] = bad[n // 2][:300] + bytes([bad[n // 2][300] ^ 4]) + bad[n // 2][301:]
for mode, thr in (("host", "1000000"), ("device", "1")):
    os.environ["BBP_DEVICE_TRANSCRIPT_MIN_BATCH"] = thr
    assert be.rangeproof_verify_batch(proofs, Vs, m, nbits, vseeds) == [0] * n
    got = be.rangeproof_verify_batch(bad, Vs, m, nbits, vseeds)
    assert [i for i, s in enumerate(got) if s] == [n // 2], got
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        be.rangeproof_verify_batch(proofs, Vs, m, nbits, vseeds)
        best = min(best, time.perf_counter() - t0)
    print(f"{mode}: {n} proofs in {best * 1e3:.2f} ms = {n / best:.0f} proofs/s", flush=True)
be.close()
