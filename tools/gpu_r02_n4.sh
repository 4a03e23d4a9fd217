#!/bin/bash
# Round-2 GPU session N4/N8b: remaining cells of the scaling table. Usage: bash tools/gpu_r02_n4.sh <N>
N=${1:-4}
mkdir -p gpurun_out
b() { name=$1; shift; timeout -k 5 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 4 "$@" > gpurun_out/r02n_bench_$name.json 2> gpurun_out/r02n_bench_$name.err; echo "bench $name exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02n_bench_$name.json').read().strip().splitlines()[-1])
    print('$name value %.1f e2e %.1f (steady %.1f) ms %.2f'%(d['value'], d['e2e']['value'], d['e2e'].get('steady_state_value') or 0, d['ms_per_step']), d.get('dp_check',{}).get('ok'), d['config'])
except Exception as e: print('parse failed', e)
PY
}
if [ "$N" = "4" ]; then b n4_default; fi
b n${N}_bins10_g64 --bins 10 --global-batch 64
