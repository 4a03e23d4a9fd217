#!/bin/bash
# Round-2 GPU session C: tests after the hf8 store fix, in-situ step traces in both parity modes, ncu launch list and
# full capture of one encoder window in f16f8 mode.
mkdir -p gpurun_out
bash tools/gpu_ci.sh tests/test_gpu_tc.py tests/test_gpu_modules.py > gpurun_out/r02c_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02c_ci.log | tail -n 10
for m in f16f8 bf16x3; do
  ESS_B200_MODE=$m timeout 300 python tools/step_trace.py gpurun_out/step_trace_r02c_$m.json > gpurun_out/step_trace_r02c_$m.txt 2>&1
  echo "trace $m exit $?"; head -n 24 gpurun_out/step_trace_r02c_$m.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 7 -c 7 -o gpurun_out/prof_window_r02c -f \
  python bench.py --mode f16f8 --steps 1 --warmup 3 --windows 3 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/prof_window_r02c.log 2>&1
echo "window capture exit $?"
timeout 300 python bench.py --mode f16f8 --steps 10 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all > gpurun_out/r02c_bench_f16f8.json 2> gpurun_out/r02c_bench_f16f8.err
echo "bench exit $?"; head -c 300 gpurun_out/r02c_bench_f16f8.json
