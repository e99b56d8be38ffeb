#!/bin/bash
# C1 (300^2) kernel durations: ncu launch list of one GMRES cycle (graph off so that ncu sees kernel launches)
mkdir -p gpurun_out
KL_USE_GRAPH=0 python bench.py --workload gmres300 --steps 95 --warmup 95 --no-extras --no-cpu-baseline > gpurun_out/r2_c1_nograph.json 2>gpurun_out/r2_c1_nograph.err || exit 1
python bench.py --workload gmres300 --steps 95 --warmup 95 --no-extras --no-cpu-baseline > gpurun_out/r2_c1_graph.json 2>>gpurun_out/r2_c1_nograph.err
KL_USE_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_c1_launches.csv python bench.py --workload gmres300 --steps 95 --warmup 95 --no-extras --no-cpu-baseline > gpurun_out/r2_c1_ncu.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r2_c1_nograph.json','gpurun_out/r2_c1_graph.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step']*1e3,2),'us/step')
PY
