#!/bin/bash
# GPU call 4 (round 2): baked-table streaming TP, fixed TMA linear, CUDA neighbour list, GPU reference baselines.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "linear" > gpurun_out/r2d_pytest_linear.log 2>&1
tail -3 gpurun_out/r2d_pytest_linear.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tp_fwd_bwd" > gpurun_out/r2d_pytest_tp.log 2>&1
tail -3 gpurun_out/r2d_pytest_tp.log
timeout 200 python tools/time_tp.py > gpurun_out/r2d_time_tp.txt 2>&1
cat gpurun_out/r2d_time_tp.txt
timeout 300 python tools/time_linear.py > gpurun_out/r2d_time_linear.txt 2>&1
cat gpurun_out/r2d_time_linear.txt
timeout 200 python tools/time_triton_ref.py > gpurun_out/r2d_time_triton_ref.txt 2>&1
tail -12 gpurun_out/r2d_time_triton_ref.txt
timeout 900 python -m pytest tests -q -m gpu --maxfail=15 > gpurun_out/r2d_pytest_all.log 2>&1
tail -12 gpurun_out/r2d_pytest_all.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err
timeout 400 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err
timeout 400 python bench.py --impl reference-gpu --steps 5 --warmup 2 > gpurun_out/r2d_bench_refgpu.json 2> gpurun_out/r2d_bench_refgpu.err
timeout 400 python bench.py --impl reference-gpu-triton --steps 5 --warmup 2 > gpurun_out/r2d_bench_refgpu_triton.json 2> gpurun_out/r2d_bench_refgpu_triton.err
cat gpurun_out/r2d_bench_refgpu.json gpurun_out/r2d_bench_refgpu_triton.json; tail -3 gpurun_out/r2d_bench_refgpu.err gpurun_out/r2d_bench_refgpu_triton.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tp_stream_kernel|linear_tma_kernel" -c 6 -o gpurun_out/r2d_prof python tools/prof_one.py > gpurun_out/r2d_ncu.log 2>&1
for f in gpurun_out/r2d_bench_c2.json gpurun_out/r2d_bench_c3.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    r=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:r[k] for k in ("value","ms_per_step","dtype") if k in r}, r.get("parity_check"), r.get("roofline"), r.get("e2e",{}).get("ms_per_step"))
    print(r.get("kernels_ms_per_step"))
except Exception as e:
    print("ERR",e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
