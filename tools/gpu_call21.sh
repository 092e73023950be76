#!/bin/bash
# GPU call 21: final state of the round -- full GPU suite, smoke, default bench line, ncu --set full of the two tensor-product backward kernels.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "baked64 or (implicit_v0 and float64)" > gpurun_out/r2u_pytest_baked64.log 2>&1; tail -4 gpurun_out/r2u_pytest_baked64.log
timeout 400 python bench.py --config c5 --dtype float64 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_c5.json 2> gpurun_out/r2u_bench_c5.err; python -c "
import json; r=json.loads(open('gpurun_out/r2u_bench_c5.json').read().strip().splitlines()[-1]); print('c5', r['ms_per_step'], r['value'], r['parity_check']); print(r['kernels_ms_per_step'])" || tail -5 gpurun_out/r2u_bench_c5.err
timeout 900 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/r2u_pytest_all.log 2>&1
tail -5 gpurun_out/r2u_pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; tail -3 gpurun_out/r2u_smoke.log
timeout 300 python bench.py > gpurun_out/r2u_bench_c2.json 2> gpurun_out/r2u_bench_c2.err
python - <<'PY'
import json
try:
    r=json.loads(open("gpurun_out/r2u_bench_c2.json").read().strip().splitlines()[-1])
    print(r["ms_per_step"], r["value"], r["e2e"]["ms_per_step"], r["parity_check"]); print(r["roofline"]); print(r["kernels_ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2u_bench_c2.err").read()[-1500:])
PY
timeout 100 python tools/time_tp.py 2>&1 | head -1 > gpurun_out/r2u_time_tp.txt; cat gpurun_out/r2u_time_tp.txt
timeout 100 python tools/exp_env2.py > gpurun_out/r2u_exp_env2.txt 2>&1; cat gpurun_out/r2u_exp_env2.txt
timeout 300 python tools/time_triton_ref.py > gpurun_out/r2u_time_triton_ref.txt 2>&1; tail -5 gpurun_out/r2u_time_triton_ref.txt
