#!/bin/bash
# Round-2 GPU session B: the f16f8 operand mode -- kernel + module parity, then bench lines in both modes.
mkdir -p gpurun_out
bash tools/gpu_ci.sh tests/test_gpu_tc.py tests/test_gpu_modules.py tests/test_golden_tc.py tests/test_gpu_kernels.py > gpurun_out/r02b_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error|T=20" gpurun_out/r02b_ci.log | tail -n 30
grep -E "^(FAILED|ERROR)" gpurun_out/test_gpu_tc.log gpurun_out/test_gpu_modules.log gpurun_out/test_golden_tc.log | head -n 40
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r02b_bench_$name.json 2> gpurun_out/r02b_bench_$name.err; echo "bench $name exit $?"; head -c 400 gpurun_out/r02b_bench_$name.json; echo; tail -n 3 gpurun_out/r02b_bench_$name.err; }
b f16f8 --mode f16f8 --steps 10 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all
b bf16x3 --mode bf16x3 --steps 10 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all
