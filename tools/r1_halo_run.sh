#!/bin/bash
# halo conv: new producer schedule + TMA store epilogue -- parity (both store paths), timing, ablation, bench batch sweep
set -u
O=gpurun_out
python -m pytest tests/test_gpu_ops.py -q -x -k "halo" > $O/t_halo_tma1.log 2>&1; echo "rc=$?" >> $O/t_halo_tma1.log
LNS_HALO_TMA=0 python -m pytest tests/test_gpu_ops.py -q -x -k "halo" > $O/t_halo_tma0.log 2>&1; echo "rc=$?" >> $O/t_halo_tma0.log
{
for tma in 0 1; do
  echo "== LNS_HALO_TMA=$tma"
  LNS_HALO_TMA=$tma python tools/ncu_conv.py 64 64 64 64 1536 1 halo
  LNS_HALO_TMA=$tma python tools/ncu_conv.py 64 64 64 128 1024 1 halo
  LNS_HALO_TMA=$tma python tools/ncu_conv.py 32 32 64 64 4096 1 halo 0 up
done
for dbg in 1 2 4 7; do echo "== TMA=1 LNS_HALO_DEBUG=$dbg"; LNS_HALO_DEBUG=$dbg python tools/ncu_conv.py 64 64 64 64 1536 1 halo; done
} > $O/halo_timing.log 2>&1
python bench.py --no-cpu-baseline --steps 5 > $O/bench_b1024.json 2> $O/bench_b1024.err
python bench.py --no-cpu-baseline --steps 3 --batch 2048 > $O/bench_b2048.json 2> $O/bench_b2048.err
python bench.py --no-cpu-baseline --steps 3 --batch 4096 > $O/bench_b4096.json 2> $O/bench_b4096.err
tail -3 $O/t_halo_tma1.log; tail -3 $O/t_halo_tma0.log; cat $O/halo_timing.log; for b in 1024 2048 4096; do cut -c1-200 $O/bench_b$b.json; done
