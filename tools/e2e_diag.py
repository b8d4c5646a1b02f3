import os, sys, time, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bbp_loader
pkg = bbp_loader.load()
gens = int(sys.argv[1]) if len(sys.argv) > 1 else 0
be = pkg.Backend(device=0, gens_capacity=gens, party_capacity=1)
n = 1 << 20
uni = hashlib.shake_256(b"x").digest(64 * n)
pts = be.from_uniform_bytes(uni)
ext, valid = be.decompress(pts)
sc = bytearray(hashlib.shake_256(b"y").digest(32 * n))
for i in range(31, 32 * n, 32):
    sc[i] &= 0x0f
h_sc = torch.frombuffer(sc, dtype=torch.uint8).pin_memory()
h_ext = torch.frombuffer(bytearray(ext), dtype=torch.uint8).pin_memory()
for k in range(4):
    t0 = time.perf_counter(); d = h_ext.cuda(non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("torch pinned H2D 128MB ms", 1e3 * (t1 - t0), flush=True)
for k in range(4):
    t0 = time.perf_counter(); r = be.msm_vartime_ptr(h_sc.data_ptr(), h_ext.data_ptr(), n); t1 = time.perf_counter()
    print("msm_vartime ms", 1e3 * (t1 - t0), flush=True)
