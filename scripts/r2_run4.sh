#!/bin/bash
# round-2 GPU run 4 (one GPU): persistent stencil CTAs, cooperative CGS2 step, one-pass GMRES step, HH reflector fusion
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -6 gpurun_out/r2_pytest4.log
for cfg in "1 0" "1 5" "1 4" "0 0"; do
  set -- $cfg
  echo "== persistent $1 occ $2"
  KL_PERSISTENT=$1 KL_PERSIST_OCC=$2 KL_SWEEP='{"2048": [[0,-1],[0,0],[64,16]], "4096": [[0,-1]], "16384": [[0,-1],[0,0]]}' python scripts/slab_sweep2.py 2>&1
done > gpurun_out/r2_slab6.log
cat gpurun_out/r2_slab6.log
for g in "1 1" "1 0" "0 0"; do
  set -- $g
  KL_COOP=$1 KL_USE_GRAPH=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-extras gmres300,hh1024,gmres4096 > gpurun_out/r2_bench_coop$1_graph$2.json 2> gpurun_out/r2_bench_coop$1_graph$2.err
  python - "$1" "$2" <<'PY'
import json,sys
f=f"gpurun_out/r2_bench_coop{sys.argv[1]}_graph{sys.argv[2]}.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), json.dumps(d["config"]["extras"]))
except Exception as e:
    print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
done
( export KRYLOV_B200_LIB=$PWD/gmres_b200/libkrylov_b200_trace.so; KL_TRACE_K1=1 python scripts/trace_cg.py 2048 > gpurun_out/r2_trace_k1_persist.log 2>&1 ); head -3 gpurun_out/r2_trace_k1_persist.log
