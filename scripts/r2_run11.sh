#!/bin/bash
# tall-skinny passes on L2-resident problems: parallel last-block reduction, CTA-count sweep (C1 300^2, C2 1024^2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -x -q > gpurun_out/r2_ts_tests.log 2>&1; tail -2 gpurun_out/r2_ts_tests.log
for tb in 0 222 148 111 74; do
  for wl in gmres300:475 hh1024:95; do
    KL_TS_BLOCKS=$tb timeout 300 python bench.py --workload ${wl%%:*} --steps ${wl##*:} --warmup ${wl##*:} --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ts_blocks $tb', d['config']['name'], round(d['value'],1),'it/s', round(d['ms_per_step']*1e3,2),'us/step')"
  done
done 2>&1 | tee gpurun_out/r2_ts_blocks.txt
