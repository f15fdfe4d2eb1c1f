# usage: bash tools/run_variants.sh [variant.so ...]  -- GPU parity tests, then bench.py per library variant
timeout -k 10 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in "" "$@"; do
  echo "== variant: ${v:-default}"
  SPLAT_B200_LIB=${v:+$PWD/$v} timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'stages', {k: round(v,3) for k,v in d['stages_ms'].items()}, 'checksum', d['frame_checksum'])"
done
if [ -f splat_b200/libsplat_b200_stats.so ]; then
  SPLAT_B200_LIB=$PWD/splat_b200/libsplat_b200_stats.so timeout -k 10 300 python tools/blend_stats.py 2>/dev/null
fi
for v in build/var/stats_*.so; do
  [ -f "$v" ] || continue
  echo "== stats: $v"
  SPLAT_B200_LIB=$PWD/$v timeout -k 10 300 python tools/blend_stats.py --frames 2 2>/dev/null
done
