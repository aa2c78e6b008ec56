"""Summarise an .ncu-rep (here, no GPU needed): key raw metrics + top stall sites of one kernel.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum [", "dram__bytes_write.sum [",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum ",
        "sm__throughput.avg.pct", "smsp__issue_active.avg.pct", "sm__inst_executed.sum ", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "smsp__inst_executed.sum ",
        "gpu__dram_throughput.avg.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg [", "smsp__cycles_active.avg ",
        "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu",
        "sm__pipe_alu_cycles_active", "sm__pipe_fma_cycles_active", "sm__pipe_fmaheavy_cycles_active"]
for vals in rows[2:]:
    name = dict(zip(hdr, vals)).get("Kernel Name", "?")
    print("== kernel:", name[:100])
    for h, u, v in zip(hdr, units, vals):
        hh = f"{h} [{u}]"
        if any(hh.startswith(k) or h == k.strip() for k in KEYS) and "per_second" not in h and "peak_sustained" not in h.replace("pct_of_peak_sustained_elapsed", ""):
            print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
data = []
for r in rows[start + 1:]:  # first kernel of the report only
    if r and r[0] == "Address":
        break
    if len(r) == len(rows[start]):
        data.append(r)
hdr = rows[start]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]]) for r in data) or 1
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"== {len(data)} SASS instructions, {tot} stall samples")
agg = {h: sum(int(r[ci[h]]) for r in data) for h in reasons}
print("   " + ", ".join(f"{h[6:]} {100 * v / tot:.1f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:topn]:
    why = {h[6:]: r[ci[h]] for h in reasons if int(r[ci[h]]) > 0.25 * max(1, int(r[ci["# Samples"]]))}
    print(f"  {100 * int(r[ci['# Samples']]) / tot:5.1f}%  exec={r[ci['Instructions Executed']]:>10}  {r[1].strip()[:70]:70s} {why}")
