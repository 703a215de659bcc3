#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
python bench.py --steps 4 --warmup 3 --workload c3s > gpurun_out/bench_c3s.json 2> gpurun_out/bench_c3s.err; tail -c 3000 gpurun_out/bench_c3s.json; tail -5 gpurun_out/bench_c3s.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
