#!/bin/bash
# usage: tools/gpu_quick.sh [pytest -k expr | none] [steps]: selected gpu tests, then bench c2 at N=1 (and N=2 when 2 GPUs are there)
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
K=${1:-none}; S=${2:-10}
if [ "$K" = all ]; then timeout 240 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
elif [ "$K" != none ]; then timeout 240 python -m pytest tests -q -m gpu --timeout 600 -x -k "$K" 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log; fi
timeout 120 python bench.py --steps $S --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $S --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; tail -3 gpurun_out/bench_c2_n2.err; fi
python tools/bench_show.py
