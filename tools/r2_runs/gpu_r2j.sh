#!/bin/bash
# device-driven two-pass near cut + PLY ingest: all GPU tests, bench (cut on / off), stress, C2/1M/10k points
tag=r2j
mkdir -p gpurun_out
out=$PWD/gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/pytest_$tag.txt
timeout -k 10 600 python bench.py --steps 50 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.log; tail -3 $out/bench_$tag.log; cut -c1-300 $out/bench_$tag.json
timeout -k 10 600 python bench.py --steps 50 --warmup 5 --near-cut 0 --no-cpu > $out/bench_nocut_$tag.json 2> $out/bench_nocut_$tag.log; cut -c1-300 $out/bench_nocut_$tag.json
: > $out/fail_$tag.txt
for i in 1 2 3 4 5 6; do
  for cfg in "--gaussians 281498 --width 1280 --height 720" "--gaussians 1000000"; do
    SPLAT_WAIT_LIMIT_S=10 timeout -k 5 120 python bench.py $cfg --steps 20 --warmup 3 --no-cpu > /tmp/o.json 2> /tmp/o.log; rc=$?
    echo "$cfg run $i rc=$rc $(python -c "import json;d=json.load(open('/tmp/o.json'));print(round(d['value'],1), d['frame_checksum'], d['near_cut']['frames_with_fallback'], d['frames_repeated'])" 2>/dev/null)" | tee -a $out/fail_$tag.txt
    if [ $rc -ne 0 ]; then grep "SplatError" /tmp/o.log | tail -1 | cut -c1-2000 >> $out/fail_$tag.txt; fi
  done
done
for n in 10000 100000; do timeout -k 5 120 python bench.py --gaussians $n --steps 50 --warmup 5 --no-cpu 2>/dev/null | cut -c1-200; done
