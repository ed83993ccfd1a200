#!/bin/bash
# Round-2 (third session) evidence on one B200 (run under gpurun).  Every ncu run follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
python bench.py > $O/bench_r02c.json 2> $O/bench_r02c.err
python bench.py --steps 3 --warmup 3 --quick > $O/bench_quick.json 2> $O/bench_quick.err &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --quick > $O/ncu_bench.log 2>&1
python tools/agg_launches.py $O/launches_bench.csv > $O/launches_bench.txt 2>&1
python tools/ncu_fablock_full.py 32 32 4736 > $O/ff_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:fablock_full2 -s 2 -c 1 -f -o $O/prof_ff2 \
    python tools/ncu_fablock_full.py 32 32 4736 > $O/ncu_ff.log 2>&1
python tools/ncu_summary.py $O/prof_ff2.ncu-rep > $O/ncu_fablock_full2_summary.txt 2>&1
LNS_TL_PREC=fp16s python tools/timeline.py ns2d 1184 20 > $O/timeline_ns2d_fp16s_c.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_c.log 2>&1
cat $O/bench_quick.json | cut -c1-300; head -20 $O/launches_bench.txt; head -24 $O/ncu_fablock_full2_summary.txt
