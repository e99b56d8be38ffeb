#!/bin/bash
# cooperative CGS2 step with persistent stage barriers and pre-issued tiles
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q -k "cooperative" > gpurun_out/r2_coop_tests.log 2>&1; tail -2 gpurun_out/r2_coop_tests.log
for coop in 0 1; do for wl in gmres300:475 gmres4096:95; do
KL_COOP=$coop timeout 200 python bench.py --workload ${wl%%:*} --steps ${wl##*:} --warmup ${wl##*:} --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('coop $coop', d['config']['name'], round(d['value'],1),'it/s', round(d['ms_per_step']*1e3,2),'us/step', (d['config'].get('parity') or {}).get('max_rel'))"
done; done
