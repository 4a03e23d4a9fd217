#!/bin/bash
# Round-2 GPU session I: row-stacked levels -- tests, then A/B bench (stacked vs dense), DDD17 bench, HALO256_WASTE A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -q -m gpu -x -k "row_stacked" --timeout 300 -p no:cacheprovider -s > gpurun_out/r02i_stacked.log 2>&1
rc=$?; echo "stacked tests exit $rc"; tail -n 25 gpurun_out/r02i_stacked.log
bash tools/gpu_ci.sh tests/test_gpu_tc.py tests/test_gpu_modules.py tests/test_golden_tc.py tests/test_gpu_uda.py > gpurun_out/r02i_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02i_ci.log | tail -n 10
grep -E "^(FAILED|ERROR)" gpurun_out/test_gpu_tc.log gpurun_out/test_gpu_modules.log | head -n 30
b() { name=$1; shift; timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all "$@" > gpurun_out/r02i_bench_$name.json 2> gpurun_out/r02i_bench_$name.err; echo "bench $name exit $?"; python tools/print_bench.py gpurun_out/r02i_bench_$name.json; tail -n 2 gpurun_out/r02i_bench_$name.err; }
b stacked
ESS_B200_STACK=0 b dense
ESS_B200_STACK=0 ESSB_TC_HALO256_WASTE=1.1 b dense_l2classic
b ddd17_stacked --workload ddd17
ESS_B200_STACK=0 b ddd17_dense --workload ddd17
