#!/bin/bash
# round-2 GPU run 3 (one GPU): full test suite after the template prune, performance check of the stencil geometry,
# chain-kernel micro-benchmark, default bench line, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -4 gpurun_out/r2_pytest3.log
for st in 1 0; do echo "== stagger $st"; KL_STENCIL_STAGGER=$st KL_SWEEP='{"2048": [[0,-1],[0,0],[64,16]], "16384": [[0,-1],[0,0],[128,32],[128,0]]}' python scripts/slab_sweep2.py 2>&1; done > gpurun_out/r2_slab5.log
cat gpurun_out/r2_slab5.log
python scripts/bench_chain.py 8192 20 > gpurun_out/r2_bench_chain.txt 2>&1; cp gpurun_out/bench_chain.json gpurun_out/r2_bench_chain.json; cat gpurun_out/r2_bench_chain.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default2.json 2> gpurun_out/r2_bench_default2.err
python scripts/show_bench.py gpurun_out/r2_bench_default2.json 2>/dev/null | head -60
# ncu: launch list of the primary bench command, then full captures of the dominant kernels
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cg16384.csv python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
python scripts/prof_kernels.py cg > gpurun_out/r2_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_stencil_tma -s 4 -c 4 -o gpurun_out/r2_prof_cg python scripts/prof_kernels.py cg > gpurun_out/r2_ncu_cg.log 2>&1
python scripts/prof_chain.py > gpurun_out/r2_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_chain_tma.*ChCheb -c 2 -o gpurun_out/r2_prof_chain python scripts/prof_chain.py > gpurun_out/r2_ncu_chain.log 2>&1
ls -la gpurun_out/*.ncu-rep
