#!/bin/bash
# quick GPU iteration: kernel + step parity tests, per-shape profile, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -15 gpurun_out/gpu_tests.log
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; tail -75 gpurun_out/profile_step.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
