"""One-line summary of a bench.py JSON line: python tools/summ.py FILE..."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    b = d.get("blindbid", {})
    g = lambda k: round(b[k]["value"]) if k in b and b[k] else None
    rp = d.get("rangeproof_m64", {})
    print(f, "N", d["n_gpus"], "msm %.0fM" % (d["value"] / 1e6), "e2e %.0fM" % (d["e2e"]["value"] / 1e6), "comp %.0fM" % (d["e2e"]["compressed_points"]["value"] / 1e6),
          "prove", g("prove"), g("prove_large_batch"), "verify", g("batch_verify"), g("batch_verify_large"), "strong", g("batch_verify_1024_total"),
          "rp", round(rp.get("prove", {}).get("value", 0)), round(rp.get("verify", {}).get("value", 0)), "check", d.get("sharded_check"))
