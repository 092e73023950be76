#!/bin/bash
# GPU call 3 (round 2): streaming TP kernels v2 (compile-time U, FFMA2) + TMA-producer tensor-core linear.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "linear" > gpurun_out/r2c_pytest_linear.log 2>&1
tail -5 gpurun_out/r2c_pytest_linear.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tp_fwd_bwd" > gpurun_out/r2c_pytest_tp.log 2>&1
tail -3 gpurun_out/r2c_pytest_tp.log
timeout 300 python tools/time_tp.py > gpurun_out/r2c_time_tp.txt 2>&1
cat gpurun_out/r2c_time_tp.txt
timeout 600 python tools/time_linear.py > gpurun_out/r2c_time_linear.txt 2>&1
cat gpurun_out/r2c_time_linear.txt
timeout 1500 python -m pytest tests -q -m gpu --maxfail=25 > gpurun_out/r2c_pytest_all.log 2>&1
tail -8 gpurun_out/r2c_pytest_all.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err
ALLEGRO_B200_FOLD_EMBED=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_c2_fold.json 2> gpurun_out/r2c_bench_c2_fold.err
timeout 900 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tp_stream_kernel|linear_tma_kernel" -c 6 -o gpurun_out/r2c_prof python tools/prof_one.py > gpurun_out/r2c_ncu.log 2>&1
for f in gpurun_out/r2c_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    r=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:r[k] for k in ("value","ms_per_step","dtype") if k in r}, r.get("parity_check"), r.get("roofline"), r.get("e2e",{}).get("ms_per_step"))
    print(r.get("kernels_ms_per_step"))
except Exception as e:
    print("ERR",e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
