#!/bin/bash
# after the no-round-trip restructure + float mode: all GPU tests, bench, then the near-cut failure with the trace build
tag=r2e
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_$tag.txt
timeout -k 10 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; tail -3 gpurun_out/bench_$tag.log; cut -c1-400 gpurun_out/bench_$tag.json
: > gpurun_out/fail_$tag.txt
for i in $(seq 1 8); do
  SPLAT_B200_LIB=$PWD/splat_b200/libsplat_b200_wdtrace.so SPLAT_WAIT_LIMIT_S=10 timeout -k 5 120 python bench.py --gaussians 281498 --width 1280 --height 720 --near-cut -1 --steps 20 --warmup 3 --no-cpu > /tmp/o.json 2> /tmp/o.log; rc=$?
  echo "C2 trace run $i rc=$rc" | tee -a gpurun_out/fail_$tag.txt
  if [ $rc -ne 0 ]; then grep "SplatError" /tmp/o.log | tail -1 >> gpurun_out/fail_$tag.txt; fi
done
grep -c watchdog gpurun_out/fail_$tag.txt
