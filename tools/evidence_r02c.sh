#!/bin/bash
# Round-2 (third session) evidence on one B200 (run under gpurun).  Every ncu run follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/t_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_c.log 2>&1
python bench.py > $O/bench_r02c.json 2> $O/bench_r02c.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_c.json 2> $O/bench_reference_c.err
python bench.py --steps 3 --warmup 3 --quick > $O/bench_quick.json 2> $O/bench_quick.err &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --quick > $O/ncu_bench.log 2>&1
python tools/agg_launches.py $O/launches_bench.csv > $O/launches_bench.txt 2>&1
LNS_TL_PREC=fp16s python tools/timeline.py ns2d 1184 20 > $O/timeline_ns2d_fp16s_c.txt 2>&1
python tools/sass_histogram.py > $O/sass_histogram.txt 2>&1
cat $O/t_gpu_all.log; tail -2 $O/smoke_c.log; cut -c1-200 $O/bench_r02c.json; head -12 $O/launches_bench.txt
