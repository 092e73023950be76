#!/bin/bash
# GPU call 17 (1 GPU): state check after the second half of round 2 -- full GPU suite, bench lines (c2 default, c3, c5), unroll
# variants of the three-warp backward, ncu launch list and --set full captures of the top kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/r2q_pytest_all.log 2>&1
tail -6 gpurun_out/r2q_pytest_all.log
timeout 300 python bench.py > gpurun_out/r2q_bench_c2.json 2> gpurun_out/r2q_bench_c2.err
python - <<'PY'
import json
try:
    r=json.loads(open("gpurun_out/r2q_bench_c2.json").read().strip().splitlines()[-1])
    print(r["ms_per_step"], r["value"], r["parity_check"]); print(r["roofline"]); print(r["kernels_ms_per_step"]); print(r["cpu_baseline"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2q_bench_c2.err").read()[-1500:])
PY
timeout 100 python tools/time_tp3.py 2>&1 | head -3 > gpurun_out/r2q_time_tp3.txt; cat gpurun_out/r2q_time_tp3.txt
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_c3.json 2> gpurun_out/r2q_bench_c3.err; python -c "
import json; r=json.loads(open('gpurun_out/r2q_bench_c3.json').read().strip().splitlines()[-1]); print('c3', r['ms_per_step'], r['value'], r['parity_check'])"
timeout 300 python bench.py --config c5 --dtype float64 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_c5.json 2> gpurun_out/r2q_bench_c5.err; python -c "
import json; r=json.loads(open('gpurun_out/r2q_bench_c5.json').read().strip().splitlines()[-1]); print('c5', r['ms_per_step'], r['value'], r['parity_check']); print(r['kernels_ms_per_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches_f32.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity-check > gpurun_out/r2q_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tp_bwd3_kernel|tp_stream_kernel|tp_stream_gyt|linear_tma_kernel|env_bwd_stream|radial_pq|env_sum_kernel|tp_smem_kernel|tp_bwd_gm" -c 14 -o gpurun_out/r2q_prof python tools/prof_one.py > gpurun_out/r2q_ncu.log 2>&1
tail -2 gpurun_out/r2q_ncu.log; ls -la gpurun_out | tail -12
