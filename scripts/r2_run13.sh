#!/bin/bash
# C1: one-pass step (chain kernel) on the small grid with short marches
mkdir -p gpurun_out
for cfg in "1048576 0" "0 0" "0 12" "0 6"; do set -- $cfg
KL_CHAIN_STEP_MIN=$1 KL_CHAIN_ROWS_MIN=$2 timeout 300 python bench.py --workload gmres300 --steps 475 --warmup 475 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step_min $1 rows_min $2', round(d['value'],1),'it/s', round(d['ms_per_step']*1e3,2),'us/step', (d['config'].get('parity') or {}).get('max_rel'))"
done
