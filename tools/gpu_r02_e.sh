#!/bin/bash
# Round-2 GPU session E: CTA-pair (cta_group::2) kernel for the N = 256 layers -- guarded first run, then A/B benches.
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "pair" --timeout 100 -p no:cacheprovider > gpurun_out/r02e_pair_tests.log 2>&1
rc=$?; echo "pair tests exit $rc"; tail -n 25 gpurun_out/r02e_pair_tests.log
if [ $rc -ne 0 ]; then echo "pair kernel not green: stopping"; nvidia-smi --query-gpu=name,utilization.gpu --format=csv; exit 0; fi
bash tools/gpu_ci.sh tests/test_gpu_tc.py tests/test_gpu_modules.py tests/test_golden_tc.py > gpurun_out/r02e_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02e_ci.log | tail -n 10
grep -E "^(FAILED|ERROR)" gpurun_out/test_gpu_tc.log gpurun_out/test_gpu_modules.log | head -n 30
b() { name=$1; shift; timeout -k 5 300 python bench.py --mode f16f8 --steps 10 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all "$@" > gpurun_out/r02e_bench_$name.json 2> gpurun_out/r02e_bench_$name.err; echo "bench $name exit $?"; python tools/print_bench.py gpurun_out/r02e_bench_$name.json; tail -n 2 gpurun_out/r02e_bench_$name.err; }
ESSB_TC_HALO256_WASTE=1.1 b pair_l01
b pair_all
ESSB_TC_PAIR=0 ESSB_TC_HALO256_WASTE=1.1 b halo_l01
ESSB_TC_HALO256_WASTE=1.1 b pair_l01_bf16x3 --mode bf16x3
