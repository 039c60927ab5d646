#!/bin/bash
# 8-GPU reproduction of the round-1 driver failure (SCALE_r01.json: rank 4 SIGABRT at N=8), with per-rank error
# reporting.  Runs 1-2 = the round-1 phase structure (no extra barriers, no parity / probe phases); run 3 = the current bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/n8
N=${1:-8}
run() {
  tag=$1; nb=$2; shift 2
  echo "=== $tag: PLIFE_BENCH_NO_BARRIERS=$nb $*" | tee -a gpurun_out/n8/summary.txt
  PLIFE_BENCH_NO_BARRIERS=$nb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --verbose "$@" > gpurun_out/n8/$tag.out 2> gpurun_out/n8/$tag.err
  echo "rc=$?" | tee -a gpurun_out/n8/summary.txt
  tail -c 1500 gpurun_out/n8/$tag.out | tee -a gpurun_out/n8/summary.txt
  grep -v "phase \|^W1\|^\*\*\*\*\|^Setting OMP" gpurun_out/n8/$tag.err | tail -n 60 | tee -a gpurun_out/n8/summary.txt
}
nvidia-smi -L | tee gpurun_out/n8/summary.txt
run r1_structure 1 --no-parity --no-probe
run r1_structure_again 1 --no-parity --no-probe
run current 0
