"""Kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`): python tools/ncu_shares.py list.csv [title]"""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1], newline="") if l.startswith('"')]
tot = defaultdict(float)
cnt = defaultdict(int)
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    ns = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] == "us":
        ns *= 1e3
    elif r["Metric Unit"] == "ms":
        ns *= 1e6
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("bbp::", "")
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
print(f"# {title}: {sum(cnt.values())} launches, GPU time {total / 1e6:.2f} ms (ncu per-launch, cold-cache and serialised)")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:34s} n={cnt[k]:4d} total_ms={tot[k] / 1e6:9.3f} share={tot[k] / total:.3f}")
