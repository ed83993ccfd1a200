"""Summarise an .ncu-rep: headline raw metrics + warp-stall samples bucketed by code region (split at barrier / MMA / TMEM markers)
    python tools/ncu_summary.py report.ncu-rep [buckets]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, val = rows[0], rows[2] if len(rows) > 2 else rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
for k in KEYS:
    for i, h in enumerate(hdr):
        if h == k:
            print(f"{k:70s} {val[i]} {rows[1][i] if len(rows) > 2 else ''}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
idx = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
S = [int(r[idx["# Samples"]]) if r[idx["# Samples"]].isdigit() else 0 for r in data]
tot = sum(S) or 1
print(f"warp-state samples: {tot} over {len(data)} SASS instructions")
MARK = ("UTCHMMA", "UTCBAR", "SYNCS", "UCGABAR", "BAR.SYNC", "LDTM", "LDGSTS", "UBLKCP", "STG", "LDG", "STS", "SHFL", "ARRIVES")
segs, prev, cur = [], 0, None
for i, r in enumerate(data):
    t = r[idx["Source"]]
    k = next((m for m in MARK if m in t), None)
    if k != cur:
        if i > prev:
            segs.append((prev, i, cur))
        prev, cur = i, k
segs.append((prev, len(data), cur))
print("region [first,last) marker  share   (regions with >= 1.5 % of the samples)")
for a, b, k in segs:
    sh = sum(S[a:b]) / tot
    if sh >= 0.015:
        print(f"  [{a:5d},{b:5d}) {str(k):9s} {100 * sh:5.1f}%   e.g. {data[a][idx['Source']].strip()[:60]}")
