#!/bin/bash
# launch list of near-cut frames (device-driven second pass) + A/B of the TMA gather in the blend + C2 point
tag=r2k
mkdir -p gpurun_out
out=$PWD/gpurun_out
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu > $out/ncu_bench_$tag.log 2>&1
for v in "" "SPLAT_B200_LIB=$PWD/splat_b200/libsplat_var_notma.so"; do
  for cut in -1 0; do
    env $v timeout -k 10 300 python bench.py --steps 50 --warmup 5 --no-cpu --near-cut $cut 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=${v:-default} cut=$cut', round(d['value'],1), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['frame_checksum'], d['near_cut']['frames_with_fallback'], d['frames_repeated'])" | tee -a $out/ab_tma_$tag.txt
  done
done
for cfg in "--gaussians 281498 --width 1280 --height 720" "--gaussians 1000000"; do
  timeout -k 5 120 python bench.py $cfg --steps 50 --warmup 5 --no-cpu 2>/tmp/e.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['frame_checksum'], d['near_cut']['frames_with_fallback'], d['frames_repeated'])" || tail -2 /tmp/e.log
done
timeout -k 10 300 python -m pytest tests/test_ply.py -m gpu -x -q 2>&1 | tail -3
