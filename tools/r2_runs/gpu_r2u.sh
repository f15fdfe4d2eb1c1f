#!/bin/bash
tag=r2u
mkdir -p gpurun_out
out=$PWD/gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/pytest_$tag.txt
for gk in "2 0" "4 1" "8 3" "8 0"; do python tools/stripe_profile.py $gk 20 2>/dev/null | tee -a $out/stripe_stages_$tag.jsonl; done
