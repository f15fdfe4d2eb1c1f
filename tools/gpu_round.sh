#!/bin/bash
# usage (under gpurun): bash tools/gpu_round.sh <tag>   -- parity tests, bench, launch list, one full ncu capture of a frame's kernels
tag=${1:-rX}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_$tag.txt
timeout -k 10 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; tail -3 gpurun_out/bench_$tag.log; cat gpurun_out/bench_$tag.json
if [ -f splat_b200/libsplat_b200_stats.so ]; then
  SPLAT_B200_LIB=$PWD/splat_b200/libsplat_b200_stats.so timeout -k 10 300 python tools/blend_stats.py 2>/dev/null | tee gpurun_out/stats_$tag.jsonl
fi
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_bench_$tag.log 2>&1
# one steady-state frame: 17 launches of the per-frame kernels (1 project, 6 hist, 6 scatter, count, emit, ranges, blend)
timeout -k 10 900 ncu --set full --clock-control none --import-source on \
    -k regex:"blend_kernel|rs_scatter|rs_hist|emit_instances|tile_count|tile_ranges|project_kernel" -s 68 -c 17 -o gpurun_out/frame_$tag -f \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out | tail -12
if [ "${CONFIGS:-0}" = "1" ]; then
  timeout -k 10 900 python tools/configs_report.py 2> gpurun_out/configs_$tag.log | tee gpurun_out/configs_$tag.jsonl | cut -c1-400
  for n in 100000 1000000; do
    timeout -k 10 300 python bench.py --gaussians $n --steps 20 --warmup 3 2>/dev/null >> gpurun_out/bench_sweep_cpu_$tag.jsonl
  done
fi
