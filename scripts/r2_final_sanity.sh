#!/bin/bash
# last run of the round: the committed library, full GPU suite + smoke + default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_sanity_tests.log 2>&1; tail -2 gpurun_out/r2_sanity_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_sanity_bench.json 2> gpurun_out/r2_sanity_bench.err; python scripts/show_bench.py gpurun_out/r2_sanity_bench.json | grep -v "^    " | cut -c1-120
