#!/bin/bash
# GPU call 6 (round 2): streaming env_bwd, thread-per-edge radial adjoint; full suite; bench lines; ncu launch list + full captures.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_zy_gpu_kernel_spec.py -q -m gpu -x -k "env_sum or radial" > gpurun_out/r2f_pytest_new.log 2>&1
tail -4 gpurun_out/r2f_pytest_new.log
timeout 900 python -m pytest tests -q -m gpu --maxfail=15 > gpurun_out/r2f_pytest_all.log 2>&1
tail -8 gpurun_out/r2f_pytest_all.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err
timeout 400 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_f32.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity-check > gpurun_out/r2f_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tp_stream_kernel|linear_tma_kernel|env_bwd_stream|radial_pq|env_sum_kernel|tp_smem_kernel|tp_bwd_gm" -c 14 -o gpurun_out/r2f_prof python tools/prof_one.py > gpurun_out/r2f_ncu.log 2>&1
for f in gpurun_out/r2f_bench_c2.json gpurun_out/r2f_bench_c3.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    r=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:r[k] for k in ("value","ms_per_step","dtype") if k in r}, r.get("parity_check"), r.get("roofline"), r.get("e2e",{}).get("ms_per_step"), r.get("cpu_baseline"))
    print(r.get("kernels_ms_per_step"))
except Exception as e:
    print("ERR",e); print(open(sys.argv[1].replace(".json",".err")).read()[-1200:])
PY
done
