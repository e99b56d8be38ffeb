#!/bin/bash
# round-2 GPU run 15: ncu --set full with SASS-level counts for the degree-4 Chebyshev chain kernel, both level lags
# (a report with sources is ~26 MB and gpurun_out/ returns at most 64 MiB: two reports per call)
mkdir -p gpurun_out
k=${1:-4}
python scripts/prof_chain2.py $k > gpurun_out/r2_prof_chain_k$k.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_tma --launch-skip 2 -c 1 -f -o gpurun_out/r2_chain_k${k}_lag2 python scripts/prof_chain2.py $k > gpurun_out/r2_ncu_chain_k${k}_lag2.log 2>&1
KRYLOV_B200_LIB=$PWD/gmres_b200/libkrylov_b200_lag1.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_tma --launch-skip 2 -c 1 -f -o gpurun_out/r2_chain_k${k}_lag1 python scripts/prof_chain2.py $k > gpurun_out/r2_ncu_chain_k${k}_lag1.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
