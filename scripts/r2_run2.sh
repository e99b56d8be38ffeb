#!/bin/bash
# round-2 GPU run 2 (one GPU): full GPU test suite, staggered-tile sweep + time line, graph effect on the
# launch-bound configs, committed bench histories, default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
for st in 1 0; do echo "== stagger $st"; KL_STENCIL_STAGGER=$st KL_SWEEP='{"2048": [[0,-1],[0,0],[48,12],[64,16],[40,10]], "4096": [[0,-1],[64,16]], "16384": [[0,-1],[128,32],[128,0],[96,24]]}' python scripts/slab_sweep2.py 2>&1; done > gpurun_out/r2_slab4.log
cat gpurun_out/r2_slab4.log
( export KRYLOV_B200_LIB=$PWD/gmres_b200/libkrylov_b200_trace.so; KL_TRACE_K1=1 python scripts/trace_cg.py 2048 > gpurun_out/r2_trace_k1_stag.log 2>&1 )
head -3 gpurun_out/r2_trace_k1_stag.log
for g in 1 0; do
  KL_USE_GRAPH=$g python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-extras gmres300,hh1024 > gpurun_out/r2_bench_graph$g.json 2> gpurun_out/r2_bench_graph$g.err
done
python scripts/make_bench_history.py > gpurun_out/r2_bench_history.log 2>&1; tail -12 gpurun_out/r2_bench_history.log
cp gpurun_out/bench_history.json tests/golden/bench_history.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_graph1.json", "gpurun_out/r2_bench_graph0.json", "gpurun_out/r2_bench_default.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "it/s e2e", d["e2e"]["value"], d["e2e"].get("value_pageable_host"), "clocks", d["clocks"], "parity", (d["config"].get("parity") or {}).get("max_rel"))
        print("   extras", json.dumps(d["config"].get("extras")))
        print("   cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
