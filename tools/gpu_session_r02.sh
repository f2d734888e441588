#!/bin/bash
# Round-2 gpurun session: tests, smoke, bench lines, ncu captures.  Outputs -> gpurun_out/
#   tools/gpu_session_r02.sh [tests] [bench] [frontend] [ncu_gl] [ncu_frontend] [ncu_mel] [launches] [small]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
  case $what in
    tests)
      timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s --durations=8 > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
      grep -E "config[245]|passed|failed|error" gpurun_out/tests.log | tail -20; tail -3 gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
      tail -2 gpurun_out/bench.log; tail -3 gpurun_out/bench.err ;;
    refarm)
      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.log ;;
    frontend)
      timeout 900 python bench.py --workload frontend --steps 10 --warmup 3 > gpurun_out/bench_frontend.log 2> gpurun_out/bench_frontend.err; echo "exit $?" >> gpurun_out/bench_frontend.err
      tail -1 gpurun_out/bench_frontend.log; tail -3 gpurun_out/bench_frontend.err ;;
    sharded1)
      timeout 900 python bench.py --workload gl_sharded --steps 3 --warmup 3 > gpurun_out/bench_sharded1.log 2> gpurun_out/bench_sharded1.err; echo "exit $?" >> gpurun_out/bench_sharded1.err
      tail -1 gpurun_out/bench_sharded1.log; tail -3 gpurun_out/bench_sharded1.err ;;
    ncu_gl)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gl_pass -s 3 -c 1 -f -o gpurun_out/glpass python tools/profile_gl.py 4 > gpurun_out/ncu_gl.log 2>&1
      tail -2 gpurun_out/ncu_gl.log ;;
    ncu_frontend)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fbank_fast -s 2 -c 1 -f -o gpurun_out/fbank16 python tools/profile_fbank.py 16000 400 > gpurun_out/ncu_fbank16.log 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fbank_fast -s 2 -c 1 -f -o gpurun_out/fbank8 python tools/profile_fbank.py 8000 400 > gpurun_out/ncu_fbank8.log 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_logmel_fast -s 2 -c 1 -f -o gpurun_out/logmel python tools/profile_logmel.py 400 > gpurun_out/ncu_logmel.log 2>&1
      tail -2 gpurun_out/ncu_logmel.log ;;
    ncu_mel)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inverse_mel_tc -s 2 -c 1 -f -o gpurun_out/invmel python tools/profile_mel.py > gpurun_out/ncu_mel.log 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mel_project_tc -s 2 -c 1 -f -o gpurun_out/melproj python tools/profile_mel.py >> gpurun_out/ncu_mel.log 2>&1
      tail -2 gpurun_out/ncu_mel.log ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv env BENCH_NO_EXTRAS=1 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1 ;;
    ncu_frames)
      # the frame-parallel kernel (cooperative, neighbour flags): one utterance of 500 frames (latency regime, 4 warps per
      # SM) and 32 x 230 frames (throughput regime, 16 warps per SM, 3 frames per warp)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gl_frames -c 1 -f -o gpurun_out/glframes500 python tools/frames_prof.py 500 > gpurun_out/ncu_frames.log 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gl_frames -c 1 -f -o gpurun_out/glframes7360 python tools/frames_prof.py 7360 >> gpurun_out/ncu_frames.log 2>&1
      tail -2 gpurun_out/ncu_frames.log ;;
    small)
      timeout 300 python tools/time_small.py > gpurun_out/small.log 2>&1; tail -12 gpurun_out/small.log ;;
  esac
done
