#!/bin/bash
# chain kernel: lines per CTA sweep at 8192^2 (degrees 4..6 and the continuation chunks of 8 / 12)
mkdir -p gpurun_out
for rows in 96 128 160 192 224 256 320; do echo "rows $rows"; KL_STENCIL_ROWS=$rows timeout 600 python scripts/bench_chain.py 8192 10 2>&1 | cut -c1-70 | sed -n '4,8p'; done > gpurun_out/r2_chain_rows2.txt 2>&1
cat gpurun_out/r2_chain_rows2.txt
