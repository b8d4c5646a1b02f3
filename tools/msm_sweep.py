"""BASELINE config 2: standalone Ristretto255 MSM sweep 2^10 .. 2^20 on one B200 (resident bases, device scalars,
CUDA-event timing, median of 20 after 3 warm-ups). Writes one JSON line per size; results are parity-checked against
the oracle by tests/test_gpu_msm.py at the same sizes."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bbp_loader
pkg = bbp_loader.load()
be = pkg.Backend(device=0, gens_capacity=0)
stream = torch.cuda.ExternalStream(be.stream(), device=0)
nmax = 1 << 20
uni = hashlib.shake_256(b"bbp-bench-points" + (0).to_bytes(8, "little")).digest(64 * nmax)
pts = be.from_uniform_bytes(uni)
raw = bytearray(hashlib.shake_256(b"bbp-bench-scalars" + (0).to_bytes(8, "little")).digest(32 * nmax))
for i in range(31, 32 * nmax, 32):
    raw[i] &= 0x0f
with torch.cuda.stream(stream):
    for lg in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10, (int(sys.argv[2]) if len(sys.argv) > 2 else 20) + 1):
        n = 1 << lg
        tab, ok = be.points_from_compressed(pts[:32 * n])
        d_sc = torch.frombuffer(bytearray(raw[:32 * n]), dtype=torch.uint8).cuda()
        d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            be.msm_points_device(d_sc.data_ptr(), n, tab, d_out.data_ptr(), None)
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            be.msm_points_device(d_sc.data_ptr(), n, tab, d_out.data_ptr(), None)
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        be.set_profiling(1)
        be.msm_points_device(d_sc.data_ptr(), n, tab, d_out.data_ptr(), None)
        stage = [round(x, 4) for x in be.msm_stage_ms()]
        be.set_profiling(0)
        plan = pkg.Backend.msm_plan(n)
        print(json.dumps({"log2_n": lg, "n": n, "ms_median": ts[10], "ms_min": ts[0], "points_per_s": n / (ts[10] * 1e-3), "window_bits": plan["c"],
                          "windows": plan["W"], "S": plan["S"], "stage_ms": stage, "result": bytes(d_out.cpu().numpy()).hex()}), flush=True)
        tab.free()
