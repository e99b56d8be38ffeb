#!/bin/bash
# chain kernel after the ring-pointer generalisation: bit-identity tests + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_chain_tests.log 2>&1; tail -2 gpurun_out/r2_chain_tests.log
timeout 600 python scripts/bench_chain.py 8192 20 > gpurun_out/r2_chain_default.txt 2>&1
cut -c1-110 gpurun_out/r2_chain_default.txt
