"""Do independent 2^20-point MSMs overlap when they alternate between contexts of one GPU? (development aid)
python tools/msm_pipeline.py [contexts] [steps]"""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bbp_loader
pkg = bbp_loader.load()
from bench import shake
C = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n = 1 << 20
bes = [pkg.Backend(device=0, gens_capacity=2048, party_capacity=1) for _ in range(C)]
uni = shake(b"bbp-bench-points" + (1000).to_bytes(8, "little"), 64 * n)
pts_c = bes[0].from_uniform_bytes(uni)
ext, valid = bes[0].decompress(pts_c)
tables = [b.points_from_extended(ext) for b in bes]
raw = bytearray(shake(b"bbp-bench-scalars" + (1000).to_bytes(8, "little"), 32 * n))
for i in range(31, 32 * n, 32):
    raw[i] &= 0x0f
d_scalars = torch.frombuffer(raw, dtype=torch.uint8).cuda()
outs = [torch.zeros(32, dtype=torch.uint8, device="cuda") for _ in bes]
torch.cuda.synchronize()
for k in range(2 * C):
    bes[k % C].msm_points_device(d_scalars.data_ptr(), n, tables[k % C], outs[k % C].data_ptr(), None)
for b in bes:
    b.sync()
assert all(bytes(o.cpu().numpy()) == bytes(outs[0].cpu().numpy()) for o in outs)
for rep in range(3):
    t0 = time.perf_counter()
    for k in range(K):
        bes[k % C].msm_points_device(d_scalars.data_ptr(), n, tables[k % C], outs[k % C].data_ptr(), None)
    for b in bes:
        b.sync()
    dt = time.perf_counter() - t0
    print(f"contexts={C} steps={K}: {1e3 * dt / K:.3f} ms per MSM, {n * K / dt / 1e6:.0f} M points/s", flush=True)
