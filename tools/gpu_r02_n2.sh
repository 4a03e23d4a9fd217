#!/bin/bash
# Round-2 GPU session N2 (2 GPUs): the 2-GPU data-parallel parity test executed on hardware + N=2 bench lines (overlapped / blocking exchange).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_dp.py -q -m gpu -s --timeout 500 -p no:cacheprovider > gpurun_out/r02n_test_gpu_dp.log 2>&1
echo "dp test exit $?"; tail -n 15 gpurun_out/r02n_test_gpu_dp.log
b() { name=$1; shift; timeout -k 5 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 4 "$@" > gpurun_out/r02n_bench_$name.json 2> gpurun_out/r02n_bench_$name.err; echo "bench $name exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02n_bench_$name.json').read().strip().splitlines()[-1])
    print('$name value %.1f e2e %.1f ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d.get('dp_check'), d['impl_config'].get('grad_exchange'))
except Exception as e: print('parse failed', e)
PY
tail -n 2 gpurun_out/r02n_bench_$name.err; }
b n2_overlap
b n2_blocking --no-overlap
b n2_bins10 --bins 10 --global-batch 64
