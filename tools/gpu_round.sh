#!/bin/bash
# round check: gpu tests, smoke, bench (c2, c3s, reference arm), then the ncu launch list of the default bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
python bench.py --steps 4 --warmup 3 --workload c3s > gpurun_out/bench_c3s.json 2> gpurun_out/bench_c3s.err; tail -3 gpurun_out/bench_c3s.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err
python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_run.log 2>&1
wc -l gpurun_out/launches.csv
