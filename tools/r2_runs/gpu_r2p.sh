#!/bin/bash
# (A) does -rdc=true tax the kernels?  same source, CDP build vs -DSPLAT_CDP=0 build without -rdc.  (B) launch list of 1M near-cut frames
tag=r2p
mkdir -p gpurun_out
out=$PWD/gpurun_out
show () { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['frame_checksum'], 'fallbacks', d['near_cut']['frames_with_fallback'], 'repeats', d['frames_repeated'])"; }
for rep in 1 2; do
for v in "" "SPLAT_B200_LIB=$PWD/splat_b200/libsplat_var_nocdp.so"; do
  for cut in -1 0; do
    env $v timeout -k 10 300 python bench.py --steps 50 --warmup 5 --no-cpu --near-cut $cut 2>/dev/null | show "lib=${v:-cdp} cut=$cut" | tee -a $out/ab_rdc_$tag.txt
  done
done
done
env SPLAT_B200_LIB=$PWD/splat_b200/libsplat_var_nocdp.so timeout -k 10 300 python bench.py --gaussians 1000000 --steps 50 --warmup 5 --no-cpu 2>/dev/null | show "nocdp 1M" | tee -a $out/ab_rdc_$tag.txt
env SPLAT_B200_LIB=$PWD/splat_b200/libsplat_var_nocdp.so timeout -k 10 300 python bench.py --gaussians 281498 --width 1280 --height 720 --steps 50 --warmup 5 --no-cpu 2>/dev/null | show "nocdp C2" | tee -a $out/ab_rdc_$tag.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches_1m_$tag.csv \
    python bench.py --gaussians 1000000 --steps 6 --warmup 3 --no-cpu > $out/ncu_bench_1m_$tag.log 2>&1
