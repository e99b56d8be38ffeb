#!/bin/bash
# round-2 final 2-GPU correctness run: multi vs single GPU (42 checks) + GMRES / BiCGSTAB / PCG bench lines with parity
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522"
timeout 900 $TR scripts/mgpu_check.py > gpurun_out/r2_final_mgpu2.log 2>&1; echo "mgpu rc=$?" >> gpurun_out/r2_final_mgpu2.log
grep -c "\[ok\]" gpurun_out/r2_final_mgpu2.log; grep "FAIL\|MGPU_CHECK\|rc=" gpurun_out/r2_final_mgpu2.log
for wl in gmres4096:95:95 bicgstab8192:20:5 pcg16384:20:5; do IFS=: read name st wu <<< "$wl"
timeout 600 $TR bench.py --gpus 2 --steps $st --warmup $wu --workload $name --no-extras --no-cpu-baseline 2> gpurun_out/r2_final_n2_$name.err | tee gpurun_out/r2_final_n2_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['name'], 'N=2', round(d['value'],1),'it/s', round(d['roofline_iter']['frac'],3), 'parity', (d['config'].get('parity') or {}).get('max_rel'))"
done
