#!/bin/bash
# Round-2 GPU session N8 (8 GPUs): where the 8-rank step time goes + one N=8 bench line.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/dp_trace.py gpurun_out/r02n_dp_trace_n8.json > gpurun_out/r02n_dp_trace_n8.txt 2>&1
echo "trace exit $?"; tail -n 16 gpurun_out/r02n_dp_trace_n8.txt
timeout -k 5 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 4 > gpurun_out/r02n_bench_n8.json 2> gpurun_out/r02n_bench_n8.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads(open('gpurun_out/r02n_bench_n8.json').read().strip().splitlines()[-1])
print('n8 value %.1f e2e %.1f ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d.get('dp_check',{}).get('ok'), d['e2e'].get('h2d_gb_per_s'), d['e2e'].get('host_wall_ms_per_step'), d['clocks'])
PY
