#!/bin/bash
# GPU call 8 (round 2, re-entry): state check -- full GPU suite, default bench line, smoke.
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -q -m gpu --maxfail=15 > gpurun_out/r2h_pytest_all.log 2>&1
tail -12 gpurun_out/r2h_pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -3 gpurun_out/r2h_smoke.log
timeout 300 python bench.py > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err
tail -c 2500 gpurun_out/r2h_bench_c2.json; tail -5 gpurun_out/r2h_bench_c2.err
