"""Per-source-line instruction counts and stall samples of the first kernel in an .ncu-rep (captured with
--import-source on from a -lineinfo build), normalised per (tile, epilogue warp) of k_knn_tc.
usage: python profiles/ncu_lines.py gpurun_out/prof_tc6.ncu-rep [tile_warps]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
T = float(sys.argv[2]) if len(sys.argv) > 2 else 6553600.0   # C2: 640 units x 1280 tiles x 8 epilogue warps
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


agg, cur = {}, None
for r in rows:
    if len(r) < 10 or r[0] == "Line No":
        continue
    if r[0] != "":
        cur = r[0]
        agg.setdefault(cur, [0, 0, r[1][:90]])
        continue
    if cur:
        agg[cur][0] += num(r[7])
        agg[cur][1] += num(r[6])
tot_i = sum(v[0] for v in agg.values())
tot_s = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {tot_i}  ({tot_i / T:.1f} per tile-warp), stall samples {tot_s}")
for ln, (ex, smp, src) in sorted(agg.items(), key=lambda x: int(x[0]) if x[0].isdigit() else 0):
    if ex / T > 0.8 or smp / tot_s > 0.004:
        print(f"{ln:>4} {ex / T:7.2f} instr/tile-warp {100 * smp / tot_s:5.1f}% samples  {src}")
