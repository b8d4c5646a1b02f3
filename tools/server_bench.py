"""Throughput of the process boundary: bbp-blindbid-server under K concurrent clients, one connection per request as the
reference's clients use it (src/futures/main.rs:64-110). Frames are encoded once (prove requests from synthetic bids, verify
requests from proofs made through the C ABI), then replayed by bbp-loadgen.
python tools/server_bench.py [n_distinct=1024] [clients=256] [requests=20000]"""
import json
import os
import struct
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbp_loader
import bench

pkg = bbp_loader.load()
capi = pkg.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
clients = int(sys.argv[2]) if len(sys.argv) > 2 else 256
requests = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
L = 8
PKG = os.path.join(ROOT, "dusk-blindbidproof_b200")


def write_frames(path, frames):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(frames)))
        for fr in frames:
            f.write(struct.pack("<I", len(fr)))
            f.write(fr)


bids = [bench.synth_bid(capi, 7000 + i, L) for i in range(n)]   # the product's own host MiMC helper
scal = lambda b: b"".join(b[k] for k in ("d", "k", "y", "y_inv", "q", "z_img", "seed"))
prove_frames = [capi.wire_prove_request(scal(b), b["pub_list"], b["toggle"]) for b in bids]
be = pkg.Backend(device=0, gens_capacity=2048, party_capacity=1)
out = be.blindbid_prove_batch(bids)
assert all(o[0] == 0 for o in out)
verify_frames = [capi.wire_verify_request(capi.wire_proof_blob(o[1], o[2], o[3]), b["q"], b["z_img"], b["seed"], b["pub_list"]) for o, b in zip(out, bids)]
be.close()
tmp = tempfile.mkdtemp()
pf, vf, sock = os.path.join(tmp, "prove.frames"), os.path.join(tmp, "verify.frames"), os.path.join(tmp, "uds")
write_frames(pf, prove_frames)
write_frames(vf, verify_frames)
result = {"list_len": L, "distinct_requests": n, "clients": clients, "prove_frame_bytes": len(prove_frames[0]), "verify_frame_bytes": len(verify_frames[0])}
for window in (os.environ.get("BBP_SERVER_WINDOWS", "200").split(",")):
    srv = subprocess.Popen([os.path.join(PKG, "bbp-blindbid-server"), "-b", sock, "-l", "info", "--window-us", window], stderr=subprocess.PIPE, text=True)
    try:
        for _ in range(1200):
            if os.path.exists(sock):
                break
            assert srv.poll() is None, srv.stderr.read()
            time.sleep(0.1)
        lg = os.path.join(PKG, "bbp-loadgen")
        legs = {}
        for name, frames, total, min_reply in (("verify", vf, requests, 2), ("prove", pf, max(clients, requests // 8), 1000)):
            subprocess.run([lg, sock, frames, str(clients), str(min(total, 4 * clients)), str(min_reply)], capture_output=True, text=True)   # warm-up (tables, allocations)
            r = subprocess.run([lg, sock, frames, str(clients), str(total), str(min_reply)], capture_output=True, text=True)
            assert r.returncode == 0, (r.stdout, r.stderr)
            legs[name] = json.loads(r.stdout)
            assert legs[name]["short_replies"] == 0, legs[name]
        result[f"window_{window}us"] = legs
    finally:
        srv.terminate()
        try:
            err = srv.communicate(timeout=60)[1]
        except subprocess.TimeoutExpired:
            srv.kill()
            err = srv.communicate()[1]
        result[f"window_{window}us"]["server_log_tail"] = err.strip().splitlines()[-1:] if err else []
print(json.dumps(result))
