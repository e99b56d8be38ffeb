#!/bin/bash
# round-2 GPU run: chain kernel (per-degree level lag / lean interior path): bit-identity tests, chain bench, N=1 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py -x -q > gpurun_out/r2_chain_tests.log 2>&1; tail -3 gpurun_out/r2_chain_tests.log
timeout 600 python scripts/bench_chain.py 8192 20 > gpurun_out/r2_chain_final.txt 2>&1; cp gpurun_out/bench_chain.json gpurun_out/r2_chain_final.json
cut -c1-110 gpurun_out/r2_chain_final.txt
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_chainfinal.json 2> gpurun_out/r2_bench_chainfinal.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_chainfinal.json').read().strip().splitlines()[-1])
print(round(d['value'],1), {k:(round(v['its_per_s'],1) if isinstance(v,dict) and 'its_per_s' in v else v) for k,v in d['config'].get('extras',{}).items()})
PY
for rows in 72 144 192; do echo "rows $rows"; KL_STENCIL_ROWS=$rows timeout 600 python scripts/bench_chain.py 8192 20 2>&1 | cut -c1-70 | head -6; done > gpurun_out/r2_chain_rows.txt 2>&1
cat gpurun_out/r2_chain_rows.txt
