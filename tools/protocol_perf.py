"""Quick timing of the batched prover / verifier on one GPU (development aid; bench.py is the contract)."""
import hashlib
import sys
import time
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader

pkg = bbp_loader.load()
capi = pkg.capi
LO = 2**252 + 27742317777372353535851937790883648493


def le(x):
    return int(x).to_bytes(32, "little")


def make_bid(i, L):
    st = hashlib.shake_256(b"bbp-bid" + i.to_bytes(8, "little")).digest(64 * (3 + L) + 8)
    k = le(int.from_bytes(st[0:64], "little") % LO)
    d = le(int.from_bytes(st[64:72], "little"))
    seed = le(int.from_bytes(st[128:192], "little") % LO)
    m = capi.mimc_hash(k, le(0))
    x = capi.mimc_hash(d, m)
    y = capi.mimc_hash(seed, x)
    z_img = capi.mimc_hash(seed, m)
    yi = pow(int.from_bytes(y, "little"), LO - 2, LO)
    q = le(int.from_bytes(d, "little") * yi % LO)
    pub = [le(int.from_bytes(st[64 * (3 + j):64 * (4 + j)], "little") % LO) for j in range(L)]
    t = i % L
    pub[t] = x
    bl = hashlib.shake_256(b"bl" + i.to_bytes(8, "little")).digest(32 * (4 + L))
    bl = b"".join(le(int.from_bytes(bl[32 * j:32 * j + 32], "little") % LO) for j in range(4 + L))
    return dict(d=d, k=k, y=y, y_inv=le(yi), q=q, z_img=z_img, seed=seed, pub_list=b"".join(pub), toggle=t, blindings=bl,
                rng_seed=hashlib.sha256(b"r%d" % i).digest())


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 16, 256]
    be = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
    bids = [make_bid(i, L) for i in range(max(sizes))]
    be.blindbid_prove_batch(bids[:2])
    for B in sizes:
        t0 = time.perf_counter()
        outs = be.blindbid_prove_batch(bids[:B])
        t1 = time.perf_counter()
        assert all(o[0] == 0 for o in outs)
        items = [dict(proof=o[1], commitments=o[2], t_c=o[3], score=b["q"], z_img=b["z_img"], seed=b["seed"], pub_list=b["pub_list"],
                      rng_seed=hashlib.sha256(b"v%d" % i).digest()) for i, (b, o) in enumerate(zip(bids, outs))]
        be.blindbid_verify_each(items[:1])
        t2 = time.perf_counter()
        st = be.blindbid_verify_each(items)
        t3 = time.perf_counter()
        ok, st2 = be.blindbid_verify_batch(items, bytes(32))
        t4 = time.perf_counter()
        assert ok and not any(st)
        print(f"L={L} B={B}: prove {1e3*(t1-t0):.2f} ms ({B/(t1-t0):.1f}/s)  verify_each {1e3*(t3-t2):.2f} ms ({B/(t3-t2):.1f}/s)  "
              f"verify_batch {1e3*(t4-t3):.2f} ms ({B/(t4-t3):.1f}/s)", flush=True)


main()
