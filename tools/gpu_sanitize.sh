#!/bin/bash
# usage (under gpurun): bash tools/gpu_sanitize.sh <tag>  -- compute-sanitizer on the smoke render (SURVEY section 4 iv)
tag=${1:-rX}
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout -k 10 600 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke:|ERROR SUMMARY|RACECHECK SUMMARY|Error|Race|hazard" | head -12
done 2>&1 | tee gpurun_out/sanitizer_$tag.txt
