"""Copy the evidence of one `tools/evidence.sh` run (gpurun_out/) into profiles/ (tracked): bench lines, aggregated ncu
launch list, per-op timeline, raw ncu metrics of the roofline kernel (+ DRAM traffic JSON) and of the FABlock kernel.
    python tools/collect_profiles.py "<build description>" """
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
what = sys.argv[1] if len(sys.argv) > 1 else "final round-1 build"

lines = [f"# bench.py lines measured on a B200 ({what}): bf16, fp16, reference arm, SW, two-phase, two-phase conditional"]
for f in ("bf16", "fp16", "reference", "sw", "twophase", "twophase_cond"):
    lines += [l.strip() for l in open(os.path.join(G, f"bench_{f}.json")) if l.startswith("{")]
open(os.path.join(P, "r01_bench_final.json"), "w").write("\n".join(lines) + "\n")
for n in (2, 8):
    src = os.path.join(G, f"bench_{n}gpu.json")
    if os.path.exists(src):
        body = [l.strip() for l in open(src) if l.startswith("{")]
        open(os.path.join(P, f"r01_bench_{n}gpu.json"), "w").write("\n".join(body) + "\n")

agg = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "agg_launches.py"), os.path.join(G, "launches_bench.csv")],
                     capture_output=True, text=True).stdout
open(os.path.join(P, "r01_launches_bench.txt"), "w").write(
    f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 3000: python bench.py --steps 1 --warmup 3 --no-cpu-baseline "
    f"({what}; first 3000 launches; cold-cache serialised times: compare SHARES)\n" + agg)
tl = os.path.join(G, "tl_ns2d_final.log")
if os.path.exists(tl):
    open(os.path.join(P, "r01_timeline_ns2d.txt"), "w").write(
        f"# python tools/timeline.py ns2d 1184 20 ({what}; serial order, CUDA events around every library call of one eager rollout)\n"
        + open(tl).read())

PAT = re.compile(r"dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|launch__block_size|launch__grid_size|"
                 r"launch__registers_per_thread|launch__shared_mem_per_block_dynamic|lts__throughput.avg.pct|sm__cycles_elapsed.avg|"
                 r"sm__inst_executed_pipe_tensor|sm__pipe_tensor|sm__throughput.avg.pct|sm__warps_active.avg.pct|smsp__inst_executed.sum|"
                 r"smsp__issue_active.avg.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|smsp__pcsamp_warps_issue_stalled")


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    res = []
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        if PAT.match(h) and v not in ("0", "") and not re.search(r"\.(max|min|sum)\.pct|subpipe_hmma_cycles|_not_issued", h):
            res.append((h, u, v))
    return res


def plain(name):
    p = os.path.join(G, name)
    return open(p).read().strip() if os.path.exists(p) else ""


halo = raw_metrics(os.path.join(G, "prof_halo.ncu-rep"))
with open(os.path.join(P, "r01_ncu_conv_halo_bench_shape.txt"), "w") as fh:
    fh.write(f"# ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 python tools/ncu_conv.py 64 64 64 64 1536 1 halo ({what})\n")
    fh.write(f"# plain run of the same command: {plain('conv_plain.log')}\n")
    fh.writelines(f"{h} [{u}] = {v}\n" for h, u, v in halo)
mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
d = {h: int(float(v.replace(",", "")) * mult[u]) for h, u, v in halo if h in ("dram__bytes_read.sum", "dram__bytes_write.sum")}
json.dump({"kernel": "conv_halo_kernel<64,2,4>", "shape": "3x3 64->64 @ 64x64, batch 1536 (bench roofline shape)",
           "source": f"ncu --set full --clock-control none, profiles/r01_ncu_conv_halo_bench_shape.txt ({what})",
           "dram_bytes_read": d["dram__bytes_read.sum"], "dram_bytes_write": d["dram__bytes_write.sum"],
           "algorithmic_bytes": 1536 * 64 * 64 * 128 * 2}, open(os.path.join(P, "r01_halo_traffic.json"), "w"))
ff = raw_metrics(os.path.join(G, "prof_ff.ncu-rep"))
with open(os.path.join(P, "r01_ncu_fablock_full_v2.txt"), "w") as fh:
    fh.write(f"# ncu --set full --clock-control none --import-source on -k regex:fablock_full -s 2 -c 1 python tools/ncu_fablock_full.py 32 32 1184 ({what})\n")
    fh.write(f"# plain run of the same command: {plain('ff_plain.log')}\n")
    fh.writelines(f"{h} [{u}] = {v}\n" for h, u, v in ff)
print("profiles/ refreshed:", what)
