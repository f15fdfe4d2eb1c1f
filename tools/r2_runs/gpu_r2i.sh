#!/bin/bash
# barriers initialised once + stripe pre-pass: all GPU tests (near-cut ones included), bench, stress of the fix
tag=r2i
mkdir -p gpurun_out
out=$PWD/gpurun_out
SPLAT_TEST_NEAR_CUT=1 timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/pytest_$tag.txt
timeout -k 10 600 python bench.py --steps 20 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.log; tail -3 $out/bench_$tag.log; cut -c1-300 $out/bench_$tag.json
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --near-cut -1 --no-cpu > $out/bench_nc_$tag.json 2> $out/bench_nc_$tag.log; tail -2 $out/bench_nc_$tag.log; cut -c1-300 $out/bench_nc_$tag.json
: > $out/fail_$tag.txt
run_c2 () {  # $1 label, rest: env assignments
  label=$1; shift
  env "$@" SPLAT_WAIT_LIMIT_S=10 timeout -k 5 120 python bench.py --gaussians 281498 --width 1280 --height 720 --near-cut -1 --steps 20 --warmup 3 --no-cpu > /tmp/o.json 2> /tmp/o.log; rc=$?
  echo "$label rc=$rc" | tee -a $out/fail_$tag.txt
  if [ $rc -ne 0 ]; then grep "SplatError" /tmp/o.log | tail -1 | cut -c1-3000 >> $out/fail_$tag.txt; fi
}
echo "=== fix applied to the failing tree (V6)"
( cd .old_r2a && for i in $(seq 1 14); do run_c2 "v6 run $i" SPLAT_B200_LIB=$PWD/splat_b200/libsplat_b200_v6.so; done )
echo "=== unfixed failing tree again (control)"
( cd .old_r2a && for i in 1 2 3 4; do run_c2 "old run $i" X=1; done )
echo "=== current tree"
for i in $(seq 1 8); do run_c2 "cur run $i" X=1; done
grep -c "rc=1" $out/fail_$tag.txt; true
