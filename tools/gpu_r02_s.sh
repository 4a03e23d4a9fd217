#!/bin/bash
# Round-2 GPU session S: CTA pairs for the N = 128 f16f8 tiles -- guarded tests first, then A/B bench.
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "encoder_conv" --timeout 120 -p no:cacheprovider > gpurun_out/r02s_tc_tests.log 2>&1
rc=$?; echo "tc tests exit $rc"; tail -n 12 gpurun_out/r02s_tc_tests.log
if [ $rc -ne 0 ]; then echo "not green: stopping"; nvidia-smi --query-gpu=name,utilization.gpu --format=csv; exit 0; fi
bash tools/gpu_ci.sh tests/test_gpu_tc.py tests/test_gpu_modules.py tests/test_gpu_uda.py > gpurun_out/r02s_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02s_ci.log | tail -n 8
b() { name=$1; shift; timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all "$@" > gpurun_out/r02s_bench_$name.json 2> gpurun_out/r02s_bench_$name.err; echo "bench $name exit $?"; python tools/print_bench.py gpurun_out/r02s_bench_$name.json 2>/dev/null | head -n 6; tail -n 1 gpurun_out/r02s_bench_$name.err; }
b pair128
ESSB_TC_PAIR128=0 b halo128
python tools/halo_probe.py 2>&1 | grep "full kernels"
ESSB_TC_PAIR128=0 python tools/halo_probe.py 2>&1 | grep "full kernels"
