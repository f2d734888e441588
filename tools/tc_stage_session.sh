#!/bin/bash
# gpurun session for the tensor-stage prototype: build, run (numerics + cycles), ncu pipe metrics of the three variants
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/ubench_tc_stage tools/ubench_tc_stage.cu || exit 1
timeout 120 /tmp/ubench_tc_stage 2000 > gpurun_out/tc_stage.txt 2>&1; echo "exit $?" >> gpurun_out/tc_stage.txt
cat gpurun_out/tc_stage.txt
timeout 600 ncu --clock-control none -k regex:k_stage -s 2 -c 6 --csv --log-file gpurun_out/tc_stage_ncu.csv \
  --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio \
  /tmp/ubench_tc_stage 500 > gpurun_out/tc_stage_ncu.log 2>&1
tail -3 gpurun_out/tc_stage_ncu.log
