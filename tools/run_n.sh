#!/bin/bash
# bench.py on N GPUs of this box (torchrun as the driver launches it); logs under gpurun_out/
cd "$(dirname "$0")/.."
N=${1:-2}; shift
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
fi
echo "N=$N rc=$?"
python - <<P
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "parity_check", "kernel_ms_per_step") if k in d}, "e2e", d["e2e"]["value"])
except Exception as e:
    print("no json:", e)
P
grep -v "^W1\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -15
