#!/bin/bash
# Round-2 GPU session H: new decoder variants on the tensor-core path, BN finalize kernel (UDA), quick bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -q -m gpu -x -k "variants or fallback" --timeout 300 -p no:cacheprovider -s > gpurun_out/r02h_variants.log 2>&1
echo "variants exit $?"; tail -n 15 gpurun_out/r02h_variants.log
timeout 600 python -m pytest tests/test_gpu_uda.py -q -m gpu -x --timeout 300 -p no:cacheprovider > gpurun_out/r02h_uda.log 2>&1
echo "uda exit $?"; tail -n 5 gpurun_out/r02h_uda.log
timeout 400 python bench.py --workload uda --steps 5 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/r02h_bench_uda.json 2> gpurun_out/r02h_bench_uda.err
echo "bench uda exit $?"; python tools/print_bench.py gpurun_out/r02h_bench_uda.json; tail -n 2 gpurun_out/r02h_bench_uda.err
