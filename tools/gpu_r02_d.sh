#!/bin/bash
# Round-2 GPU session D: wide (N = 256) halo kernel A/B, enc0 operand format A/B in the f16f8 mode.
mkdir -p gpurun_out
bash tools/gpu_ci.sh tests/test_gpu_tc.py tests/test_gpu_modules.py > gpurun_out/r02d_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02d_ci.log | tail -n 10
grep -E "^(FAILED|ERROR)" gpurun_out/test_gpu_tc.log gpurun_out/test_gpu_modules.log | head -n 30
b() { name=$1; shift; timeout 600 python bench.py --mode f16f8 --steps 10 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all "$@" > gpurun_out/r02d_bench_$name.json 2> gpurun_out/r02d_bench_$name.err; echo "bench $name exit $?"; python tools/print_bench.py gpurun_out/r02d_bench_$name.json; tail -n 2 gpurun_out/r02d_bench_$name.err; }
b default
ESSB_TC_HALO256=0 b classic256
ESSB_TC_HALO256_WASTE=1.1 b halo256_l01
ESS_B200_F16F8_ENC0=hf8 b enc0_hf8
ESS_B200_MODE=f16f8 timeout 300 python tools/step_trace.py gpurun_out/step_trace_r02d_f16f8.json > gpurun_out/step_trace_r02d_f16f8.txt 2>&1
echo "trace exit $?"; head -n 12 gpurun_out/step_trace_r02d_f16f8.txt; tail -n 3 gpurun_out/step_trace_r02d_f16f8.txt
