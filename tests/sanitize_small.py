"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck): MSM (variable / fixed / skewed),
codecs, prove + verify with the device-side RNG / witness / transcript / hybrid-IPP paths forced, range proof."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # orc: the oracle is the checker here, as in the tests beside this script
os.environ["BBP_DEVICE_RNG_MIN_BATCH"] = "1"
os.environ["BBP_DEVICE_TRANSCRIPT_MIN_BATCH"] = "1"
os.environ["BBP_IPP_HYBRID"] = "2"
import bbp_loader, orc
pkg = bbp_loader.load()
be = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
n = 700
pts = be.from_uniform_bytes(hashlib.shake_256(b"san").digest(64 * n))
scs = orc.random_scalars(1, n)
assert be.msm_optional(scs, pts) == orc.msm(scs, pts, algo=1)
assert be.msm_optional(scs[:32] * n, pts) == orc.msm(scs[:32] * n, pts, algo=1)     # skewed buckets -> fold path
ext, valid = be.decompress(pts[:32 * 40]); assert be.compress(ext) == pts[:32 * 40]
bids = []
for i in range(2):
    b = orc.make_bid(9000 + i, 2, i % 2); b["blindings"] = orc.bid_blindings(9000 + i, 2); b["rng_seed"] = bytes([i]) * 32
    bids.append(b)
outs = be.blindbid_prove_batch(bids)
rc, proof, comm, tc = orc.blindbid_prove(bids[0], bids[0]["blindings"], bids[0]["rng_seed"])
assert outs[0] == (0, proof, comm, tc)
items = [dict(proof=o[1], commitments=o[2], t_c=o[3], score=b["q"], z_img=b["z_img"], seed=b["seed"], pub_list=b["pub_list"], rng_seed=bytes(32))
         for b, o in zip(bids, outs)]
assert be.blindbid_verify_each(items) == [0, 0]
ok, st = be.blindbid_verify_batch(items, bytes(32)); assert ok
bad = [dict(x) for x in items]; p = bytearray(bad[1]["proof"]); p[-1] ^= 1; bad[1]["proof"] = bytes(p)
ok, st = be.blindbid_verify_batch(bad, bytes(32)); assert not ok and st[0] == 0 and st[1] != 0
be.close()
be = pkg.Backend(device=0, gens_capacity=8, party_capacity=4)
bl = b"".join((7 + i).to_bytes(32, "little") for i in range(4))
rc, pf, V = be.rangeproof_prove([1, 2, 3, 255], bl, 8, bytes(32)); assert rc == 0
assert be.rangeproof_verify(pf, V, 8, bytes(32)) == 0
# a batch that does not fill the replay kernel's last block (5 proofs, 4 warps per block), one proof corrupted
vals = [[(3 * k + j) & 255 for j in range(4)] for k in range(5)]
st, pfs, Vs = be.rangeproof_prove_batch(vals, bl * 5, 4, 8, bytes(range(32)) * 5)
assert st == [0] * 5
assert be.rangeproof_verify_batch(pfs, Vs, 4, 8, bytes(32) * 5) == [0] * 5
pfs[3] = pfs[3][:150] + bytes([pfs[3][150] ^ 1]) + pfs[3][151:]
got = be.rangeproof_verify_batch(pfs, Vs, 4, 8, bytes(32) * 5)
assert got[3] != 0 and [got[i] for i in (0, 1, 2, 4)] == [0] * 4, got
be.close()
print("sanitize_small ok")
