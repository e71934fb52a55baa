#!/usr/bin/env python
"""Digest an `ncu --csv --metrics ...` log: per-kernel launch count, total/mean of each metric, share of
gpu__time_duration.  usage: ncu_launches.py log.csv [--traffic-json out.json kernel-substring]"""
import collections, csv, json, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.defaultdict(lambda: collections.defaultdict(int))
SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for row in csv.DictReader(lines):
    k = re.sub(r"\(kf::.*", "", row["Kernel Name"]).replace("void kf::", "")
    m, u = row["Metric Name"], row["Metric Unit"]
    v = float(row["Metric Value"].replace(",", "")) * SCALE.get(u, 1.0)
    tot[k][m] += v
    cnt[k][m] += 1
T = sum(t.get("gpu__time_duration.sum", 0.0) for t in tot.values()) or 1.0
print(f"{'kernel':46s} {'launches':>8s} {'total ms':>10s} {'share':>7s}  other metrics (mean per launch)")
for k, t in sorted(tot.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0.0)):
    ms = t.get("gpu__time_duration.sum", 0.0)
    n = max(cnt[k].values())
    extra = "  ".join(f"{m}={t[m]/cnt[k][m]:.4g}" for m in t if m != "gpu__time_duration.sum")
    print(f"{k[:46]:46s} {n:8d} {ms:10.3f} {ms/T*100:6.1f}%  {extra}")
if "--traffic-json" in sys.argv:
    i = sys.argv.index("--traffic-json")
    out, sub = sys.argv[i + 1], sys.argv[i + 2]
    for k, t in tot.items():
        if sub in k:
            n = cnt[k]["dram__bytes_read.sum"]
            b = (t["dram__bytes_read.sum"] + t["dram__bytes_write.sum"]) / n
            json.dump({"kernel": k, "launches": n, "dram_bytes_per_launch": b,
                       "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over {n} launches ({path})"},
                      open(out, "w"), indent=1)
            print("wrote", out, b)
