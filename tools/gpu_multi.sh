#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> [tests]
tag=${1:-r2m}; N=${2:-2}; tests=${3:-1}
mkdir -p gpurun_out
out=$PWD/gpurun_out
nvidia-smi -L | tee $out/smi_${tag}_$N.txt
if [ "$tests" = "1" ]; then
  timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest_multi_${tag}_$N.txt
fi
run () {  # $1 name, rest: bench args
  name=$1; shift
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 30 --warmup 5 "$@" > $out/bench_${name}_${tag}_$N.json 2> $out/bench_${name}_${tag}_$N.log
  echo "rc=$?"; grep -E "rebalance|stage ms|rows \[" $out/bench_${name}_${tag}_$N.log | tail -$((2*N+6)); cut -c1-260 $out/bench_${name}_${tag}_$N.json
}
echo "=== C3 1080p x$N"; run c3
echo "=== C4 4K x$N"; run c4 --gaussians 5800000 --width 3840 --height 2160
if [ "${SWEEP:-0}" = "1" ]; then
  for n in 100000 1000000 10000000; do echo "=== C5 N=$n x$N"; run c5_$n --gaussians $n; done
fi
