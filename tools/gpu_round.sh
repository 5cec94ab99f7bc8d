#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), and an ncu --set full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
if [ "$1" = "ncu" ]; then
  python tools/one_step.py > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc_kernel -s 300 -c 3 -o gpurun_out/prof_conv_tc python tools/one_step.py > gpurun_out/ncu_full.log 2>&1
fi
tail -3 gpurun_out/gpu_tests.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
