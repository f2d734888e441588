#!/bin/bash
# One gpurun session: tests, smoke, micro-benchmarks, bench, ncu launch list.  Outputs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 120 tools/ubench > gpurun_out/ubench.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
fi
tail -5 gpurun_out/tests.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log
