#!/bin/bash
# A/B session for the Griffin-Lim iteration kernel: parity tests, pass timing of both formulations, one ncu capture.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gl_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -5
python tools/time_pass.py 0 2>&1 | tail -1
S2ST_GL_KERNEL=classic python tools/time_pass.py 0 2>&1 | tail -1
if [ "$1" = "ncu" ]; then
  timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_gl_pass_r64 -s 3 -c 1 -f -o gpurun_out/glpass_r64 python tools/profile_gl.py 4 > gpurun_out/ncu_r64.log 2>&1; tail -2 gpurun_out/ncu_r64.log
fi
