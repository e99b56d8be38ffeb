#!/bin/bash
# round-2 GPU run 1: regression tests + PDL / push-halo experiments (N=1 and N=2)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -3 gpurun_out/r2_pytest1.log
for pdl in 1 0; do
  KL_PDL=$pdl python bench.py --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n1_pdl$pdl.json 2> gpurun_out/r2_bench_n1_pdl$pdl.err
done
KL_PDL=1 python scripts/slab_sweep.py > gpurun_out/r2_slab_pdl1.log 2>&1
KL_PDL=0 python scripts/slab_sweep.py > gpurun_out/r2_slab_pdl0.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/mgpu_check.py > gpurun_out/r2_mgpu2.log 2>&1; echo "mgpu rc=$?" >> gpurun_out/r2_mgpu2.log
tail -3 gpurun_out/r2_mgpu2.log
for cfg in "1 1" "1 0" "0 1" "0 0"; do
  set -- $cfg
  KL_PUSH_HALO=$1 KL_PDL=$2 timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n2_push$1_pdl$2.json 2> gpurun_out/r2_bench_n2_push$1_pdl$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n*_p*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), 'it/s', round(d['ms_per_step']*1e3,1),'us', 'e2e', round(d['e2e']['value'],1), [ (k['name'].split()[0], round(k['avg_us'],1)) for k in d.get('kernels',[])])
    except Exception as e: print(f,'ERR',e)
PY
