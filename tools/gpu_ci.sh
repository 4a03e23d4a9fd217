#!/bin/bash
# Runs the GPU test files one process per file (a sticky CUDA error in one file cannot poison the
# others), each under a timeout, logging into gpurun_out/.  Usage: bash tools/gpu_ci.sh [files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
FILES=${@:-"tests/test_gpu_kernels.py tests/test_gpu_tc.py tests/test_gpu_modules.py tests/test_gpu_uda.py tests/test_gpu_teacher.py tests/test_golden_tc.py tests/test_gpu_dp.py"}
rc=0
for f in $FILES; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s --timeout 600 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  r=$?
  echo "$f exit $r"; tail -n 40 gpurun_out/$name.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
