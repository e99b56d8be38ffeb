#!/bin/bash
# programmatic dependent launch across the GMRES / Householder step kernels: full GPU tests + extras
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pdl_tests.log 2>&1; tail -3 gpurun_out/r2_pdl_tests.log
for pdl in 1 0; do
KL_PDL=$pdl timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2_pdl_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('KL_PDL=$pdl', round(d['value'],1), {k:v['its_per_s'] for k,v in d['config']['extras'].items()})"
done
