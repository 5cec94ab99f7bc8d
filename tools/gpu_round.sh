#!/bin/bash
# One GPU-box session for the round's evidence: parity tests, smoke, bench (both arms), per-kernel times, the ncu launch
# list and an ncu --set full capture of the dominant kernel.   bash tools/gpu_round.sh <tag>   (outputs in gpurun_out/)
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/bench_ref.err
timeout 300 python bench.py --size 128 --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_128_b16.json 2> gpurun_out/bench_128.err
timeout 200 python tools/kernel_times.py 80 > gpurun_out/${TAG}_kernel_times.log 2>&1
python tools/one_step.py > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${TAG}.csv python tools/one_step.py > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_tc_persist_kernel -s 60 -c 3 -f -o gpurun_out/prof_${TAG}_persist python tools/one_step.py > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/gpu_tests.log; tail -n 2 gpurun_out/smoke.log; cat gpurun_out/${TAG}_bench.json; cat gpurun_out/${TAG}_bench_reference.json; cat gpurun_out/${TAG}_bench_128_b16.json | cut -c1-200; tail -n 2 gpurun_out/ncu_full.log
