#!/bin/bash
# One gpurun session: tests, smoke, bench, optional ncu captures.  Outputs -> gpurun_out/
#   tools/gpu_session.sh [tests] [bench] [ncu] [launches] [frontend]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
  case $what in
    tests)
      timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
      tail -5 gpurun_out/tests.log; tail -3 gpurun_out/smoke.log ;;
    bench)
      timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
      timeout 300 python tools/step_breakdown.py > gpurun_out/breakdown.log 2>&1
      tail -2 gpurun_out/bench.log; tail -4 gpurun_out/breakdown.log ;;
    ncu)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gl_pass -s 3 -c 1 -f -o gpurun_out/glpass python tools/profile_gl.py 4 > gpurun_out/ncu.log 2>&1
      tail -3 gpurun_out/ncu.log ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1 ;;
    frontend)
      timeout 600 python tools/bench_frontend.py > gpurun_out/frontend.log 2>&1; tail -8 gpurun_out/frontend.log ;;
  esac
done
