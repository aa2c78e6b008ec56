"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/launch_summary.py gpurun_out/launches.csv [bench.json] > profiles/rN_launches_summary.txt"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.Counter()
cnt = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "")
    tot[name] += float(r[14]) / 1e6
    cnt[name] += 1
total = sum(tot.values())
live = ""
if len(sys.argv) > 2:
    b = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    live = f"  Live share of k_knn_tc in bench.py: {b['roofline']['kernel_share_of_step']:.3f}"
print(f"# {len(rows)} launches, {total:.2f} ms summed (cold-cache, serialised: compare SHARES).{live}")
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for name, ms in tot.most_common():
    print(f"{name[-60:]:60s} {cnt[name]:8d} {ms:10.3f} {100 * ms / total:6.1f}%")
