#!/bin/bash
# Round evidence on one B200 (run under gpurun): bench lines, reference arm, ncu launch list of the bench command, ncu --set
# full captures of the roofline kernel and the two FABlock kernels.  Every ncu run follows a plain run of the same command.
set -u
O=gpurun_out
python bench.py > $O/bench_bf16.json 2> $O/bench_bf16.err
python bench.py --precision fp16 --no-cpu-baseline --steps 5 > $O/bench_fp16.json 2> $O/bench_fp16.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for w in sw twophase twophase_cond; do
  python bench.py --workload $w --steps 5 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_short.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
python tools/ncu_conv.py 64 64 64 64 1536 1 halo > $O/conv_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -f -o $O/prof_halo \
    python tools/ncu_conv.py 64 64 64 64 1536 1 halo > $O/ncu_halo.log 2>&1
python tools/ncu_fablock_full.py 32 32 1184 > $O/ff_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:fablock_full -s 2 -c 1 -f -o $O/prof_ff \
    python tools/ncu_fablock_full.py 32 32 1184 > $O/ncu_ff.log 2>&1
cat $O/bench_bf16.json | cut -c1-400
