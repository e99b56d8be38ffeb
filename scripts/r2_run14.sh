#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_chain.py tests/test_gpu_drivers.py tests/test_gpu_reference.py -x -q > gpurun_out/r2_tests14.log 2>&1; tail -3 gpurun_out/r2_tests14.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2_bench14.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), {k:(v['its_per_s'], v['roofline_iter_frac']) for k,v in d['config']['extras'].items()})"
