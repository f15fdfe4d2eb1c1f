#!/bin/bash
# round-2 first GPU pass: parity tests with the new defaults + full-size configs, bench with the parity key,
# the reference arm's timing, and the near-cut hang repro under the watchdog
tag=r2a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$tag.txt
timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_$tag.txt
timeout -k 10 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; tail -3 gpurun_out/bench_$tag.log; cut -c1-600 gpurun_out/bench_$tag.json
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/ref_$tag.json 2> gpurun_out/ref_$tag.log; tail -4 gpurun_out/ref_$tag.log
echo "--- near cut, 1M, automatic"
SPLAT_WAIT_LIMIT_S=15 timeout -k 10 150 python bench.py --gaussians 1000000 --near-cut -1 --steps 3 --warmup 3 --no-cpu > gpurun_out/nc1m_$tag.json 2> gpurun_out/nc1m_$tag.log; echo "rc=$?"; tail -5 gpurun_out/nc1m_$tag.log
echo "--- near cut, 1M, automatic, debug sync"
SPLAT_DEBUG_SYNC=1 SPLAT_WAIT_LIMIT_S=15 timeout -k 10 150 python bench.py --gaussians 1000000 --near-cut -1 --steps 3 --warmup 3 --no-cpu > gpurun_out/nc1m_dbg_$tag.json 2> gpurun_out/nc1m_dbg_$tag.log; echo "rc=$?"; tail -5 gpurun_out/nc1m_dbg_$tag.log
echo "--- near cut tests"
SPLAT_TEST_NEAR_CUT=1 SPLAT_WAIT_LIMIT_S=15 timeout -k 10 300 python -m pytest tests -m gpu -x -q -k near_cut 2>&1 | tail -8 | tee gpurun_out/pytest_nc_$tag.txt
