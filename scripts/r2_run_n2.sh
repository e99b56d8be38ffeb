#!/bin/bash
# round-2 GPU run on 2 GPUs: multi-GPU parity check + bench variants
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522"
timeout 900 $TR scripts/mgpu_check.py > gpurun_out/r2_mgpu2b.log 2>&1; echo "mgpu rc=$?" >> gpurun_out/r2_mgpu2b.log
grep -c "\[ok\]" gpurun_out/r2_mgpu2b.log; grep "FAIL\|MGPU_CHECK\|rc=" gpurun_out/r2_mgpu2b.log
for cfg in "1 1 1" "1 1 0" "0 0 0"; do
  set -- $cfg
  KL_PUSH_HALO=$1 KL_PDL=$2 KL_REVERSE=$3 timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n2b_push$1_pdl$2_rev$3.json 2> gpurun_out/r2_bench_n2b_push$1_pdl$2_rev$3.err
done
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload bicgstab8192 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n2b_bicgstab.json 2> gpurun_out/r2_bench_n2b_bicgstab.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n2b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), 'it/s', round(d['ms_per_step']*1e3,1),'us', 'parity', (d['config'].get('parity') or {}).get('max_rel'), [ (k['name'].split()[0], round(k['avg_us'],1)) for k in d.get('kernels',[])])
    except Exception as e: print(f,'ERR',e)
PY
