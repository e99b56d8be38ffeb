#!/bin/bash
# two-level grid reduction in the tall-skinny passes: tests, time line, C1 / C2 / C3 rates
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_chain.py tests/test_gpu_large.py tests/test_gpu_dense.py -x -q > gpurun_out/r2_tests16.log 2>&1; tail -3 gpurun_out/r2_tests16.log
KRYLOV_B200_LIB=$PWD/gmres_b200/libkrylov_b200_trace.so python scripts/trace_ts.py 300 95 300 48 1024 95 2>&1 | tail -18
for wl in gmres300:475 hh1024:95 gmres4096:95; do
timeout 300 python bench.py --workload ${wl%%:*} --steps ${wl##*:} --warmup ${wl##*:} --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['name'], round(d['value'],1),'it/s', round(d['ms_per_step']*1e3,2),'us/step', round(d['roofline_iter']['frac'],3), (d['config'].get('parity') or {}).get('max_rel'))"
done
