"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN_summary.txt"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms total (ncu-serialised, cold cache: compare SHARES)")
print(f"{'kernel':72s} {'n':>6s} {'total_ms':>10s} {'share':>7s} {'avg_us':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:72s} {v[0]:6d} {v[1]:10.2f} {100 * v[1] / tot:6.1f}% {v[1] / v[0] * 1000:9.1f}")
