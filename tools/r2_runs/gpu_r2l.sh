#!/bin/bash
# CDP-launched second pass: all GPU tests, bench (cut on/off), C2 / 1M / 10k / 100k points, stress
tag=r2q
mkdir -p gpurun_out
out=$PWD/gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/pytest_$tag.txt
timeout -k 10 600 python bench.py --steps 50 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.log; tail -2 $out/bench_$tag.log
show () { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['frame_checksum'], 'fallbacks', d['near_cut']['frames_with_fallback'], 'repeats', d['frames_repeated'], 'launches', d['gpu_launches'])"; }
show C3 < $out/bench_$tag.json
timeout -k 10 600 python bench.py --steps 50 --warmup 5 --near-cut 0 --no-cpu 2>/dev/null | show C3-nocut
for cfg in "--gaussians 281498 --width 1280 --height 720" "--gaussians 1000000" "--gaussians 10000" "--gaussians 100000" "--gaussians 10000000"; do
  SPLAT_WAIT_LIMIT_S=10 timeout -k 5 200 python bench.py $cfg --steps 50 --warmup 5 --no-cpu 2>/tmp/e.log | show "$cfg" || tail -2 /tmp/e.log
done
for i in 1 2 3 4 5 6; do
  SPLAT_WAIT_LIMIT_S=10 timeout -k 5 120 python bench.py --gaussians 281498 --width 1280 --height 720 --steps 20 --warmup 3 --no-cpu > /tmp/o.json 2> /tmp/o.log; echo "C2 stress $i rc=$?"
done
