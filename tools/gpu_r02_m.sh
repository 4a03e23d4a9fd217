#!/bin/bash
# Round-2 GPU session M: is the end-to-end number stable when bench.py is the FIRST process on a fresh box (what the driver does)?
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout -k 5 600 python bench.py --no-torch-gpu-baseline --no-cpu-baseline > gpurun_out/r02m_bench_first_$i.json 2> gpurun_out/r02m_bench_first_$i.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02m_bench_first_$i.json').read().strip().splitlines()[-1])
print('run $i value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d['e2e'].get('host_wall_ms_per_step'), 'h2d', round(d['e2e']['h2d_gb_per_s'],1), d['clocks'])
PY
done
