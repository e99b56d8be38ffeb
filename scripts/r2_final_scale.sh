#!/bin/bash
# round-2 final strong-scaling run on one 8-GPU box: CG 16384^2 at N = 8, 4, 2, 1 (the driver's launch line), GMRES 4096^2 at N = 8
mkdir -p gpurun_out
run() { # N port extra...
  local n=$1 port=$2; shift 2
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 "$@"
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; fi
}
for n in 8 4 2 1; do
  run $n $((29600+n)) --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_final_cg16384_n$n.json 2> gpurun_out/r2_final_cg16384_n$n.err
done
run 8 29611 --steps 95 --warmup 95 --workload gmres4096 --no-extras --no-cpu-baseline > gpurun_out/r2_final_gmres4096_n8.json 2> gpurun_out/r2_final_gmres4096_n8.err
run 8 29612 --steps 20 --warmup 5 --workload pcg16384 --no-extras --no-cpu-baseline > gpurun_out/r2_final_pcg16384_n8.json 2> gpurun_out/r2_final_pcg16384_n8.err
python - <<'PY'
import json,glob
base=None
for f in ['gpurun_out/r2_final_cg16384_n%d.json'%n for n in (1,2,4,8)]+['gpurun_out/r2_final_gmres4096_n8.json','gpurun_out/r2_final_pcg16384_n8.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if base is None: base=d['value']
        print(f.split('/')[-1], round(d['value'],1), 'it/s', round(d['ms_per_step']*1e3,1),'us', 'x%.2f'%(d['value']/base), 'roofline_iter', round(d['roofline_iter']['frac'],3), 'parity', (d['config'].get('parity') or {}).get('max_rel'), d.get('clocks',{}).get('sm_mhz'))
    except Exception as e: print(f,'ERR',e)
PY
