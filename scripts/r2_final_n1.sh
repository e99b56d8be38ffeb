#!/bin/bash
# round-2 final single-GPU validation: full GPU test suite, smoke, default bench, reference arm, chain bench, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1; tail -3 gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; tail -c 300 gpurun_out/r2_final_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; tail -c 200 gpurun_out/r2_final_bench_ref.json
timeout 600 python scripts/bench_chain.py 8192 20 > gpurun_out/r2_chain_final.txt 2>&1; cp gpurun_out/bench_chain.json gpurun_out/r2_chain_final.json; cut -c1-110 gpurun_out/r2_chain_final.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2_final_ncu_launch.log 2>&1
du -sh gpurun_out
