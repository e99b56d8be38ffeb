#!/bin/bash
# round-2 GPU run on 8 GPUs: CG 16384^2 strong scaling with the fused halo push / PDL / reverse-march variants
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
for cfg in "1 1 1" "1 1 0" "0 0 0"; do
  set -- $cfg
  KL_PUSH_HALO=$1 KL_PDL=$2 KL_REVERSE=$3 timeout 900 $TR bench.py --gpus 8 --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n8_push$1_pdl$2_rev$3.json 2> gpurun_out/r2_bench_n8_push$1_pdl$2_rev$3.err
done
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n8_driver.json 2> gpurun_out/r2_bench_n8_driver.err
timeout 600 $TR bench.py --gpus 8 --steps 95 --warmup 95 --workload gmres4096 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n8_gmres4096.json 2> gpurun_out/r2_bench_n8_gmres4096.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), 'it/s', round(d['ms_per_step']*1e3,1),'us', 'roofline_iter', round(d['roofline_iter']['frac'],3), 'parity', (d['config'].get('parity') or {}).get('max_rel'), [ (k['name'].split()[0], round(k['avg_us'],1)) for k in d.get('kernels',[])])
    except Exception as e: print(f,'ERR',e)
PY
