#!/bin/bash
# round-2 GPU run 5 (one GPU): balanced persistent grid vs one CTA per tile; new tests; C1 bench
mkdir -p gpurun_out
for cfg in "0 0" "1 0" "1 5"; do
  set -- $cfg
  echo "== persistent $1 occ $2"
  KL_PERSISTENT=$1 KL_PERSIST_OCC=$2 KL_SWEEP='{"2048": [[0,-1],[0,0],[64,16]], "4096": [[0,-1]], "16384": [[0,-1],[0,0]]}' python scripts/slab_sweep2.py 2>&1
done > gpurun_out/r2_slab7.log
cat gpurun_out/r2_slab7.log
( export KRYLOV_B200_LIB=$PWD/gmres_b200/libkrylov_b200_trace.so; KL_PERSISTENT=1 KL_TRACE_K1=1 python scripts/trace_cg.py 2048 > gpurun_out/r2_trace_k1_persist_bal.log 2>&1 ); head -3 gpurun_out/r2_trace_k1_persist_bal.log
python -m pytest tests/test_gpu_chain.py tests/test_gpu_large.py -m gpu -q > gpurun_out/r2_pytest7.log 2>&1; tail -5 gpurun_out/r2_pytest7.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-extras gmres300,hh1024 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_c1.json').read().strip().splitlines()[-1]); print(round(d['value'],1), json.dumps(d['config']['extras']))"
