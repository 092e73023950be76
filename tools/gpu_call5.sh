#!/bin/bash
# GPU call 5 (round 2): radial_pq + W1 fold, slice-fit fix for wide linears, CUDA neighbour list tests, GPU reference baselines.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zy_gpu_kernel_spec.py tests/test_zv_gpu_nlist.py tests/test_zzz_gpu_fold_embed.py -q -m gpu > gpurun_out/r2e_pytest_new.log 2>&1
tail -6 gpurun_out/r2e_pytest_new.log
timeout 200 python tools/time_tp.py > gpurun_out/r2e_time_tp.txt 2>&1
head -2 gpurun_out/r2e_time_tp.txt
timeout 300 python tools/time_linear.py > gpurun_out/r2e_time_linear.txt 2>&1
head -12 gpurun_out/r2e_time_linear.txt
timeout 200 python tools/time_triton_ref.py > gpurun_out/r2e_time_triton_ref.txt 2>&1
tail -9 gpurun_out/r2e_time_triton_ref.txt
timeout 900 python -m pytest tests -q -m gpu --maxfail=15 > gpurun_out/r2e_pytest_all.log 2>&1
tail -8 gpurun_out/r2e_pytest_all.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r2e_bench_c2.json 2> gpurun_out/r2e_bench_c2.err
ALLEGRO_B200_PLAIN_BWD=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench_c2_plainbwd.json 2> gpurun_out/r2e_bench_c2_plainbwd.err
ALLEGRO_B200_FOLD_RADIAL=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench_c2_nofoldradial.json 2> gpurun_out/r2e_bench_c2_nofoldradial.err
timeout 400 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_c3.json 2> gpurun_out/r2e_bench_c3.err
timeout 400 python bench.py --config c5 --dtype float64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err
timeout 400 python bench.py --impl reference-gpu --steps 5 --warmup 2 > gpurun_out/r2e_bench_refgpu.json 2> gpurun_out/r2e_bench_refgpu.err
timeout 400 python bench.py --impl reference-gpu-triton --steps 5 --warmup 2 > gpurun_out/r2e_bench_refgpu_triton.json 2> gpurun_out/r2e_bench_refgpu_triton.err
cat gpurun_out/r2e_bench_refgpu.json gpurun_out/r2e_bench_refgpu_triton.json; tail -n 3 gpurun_out/r2e_bench_refgpu.err gpurun_out/r2e_bench_refgpu_triton.err
for f in gpurun_out/r2e_bench_c2.json gpurun_out/r2e_bench_c2_plainbwd.json gpurun_out/r2e_bench_c2_nofoldradial.json gpurun_out/r2e_bench_c3.json gpurun_out/r2e_bench_c5.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    r=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:r[k] for k in ("value","ms_per_step","dtype") if k in r}, r.get("parity_check"), r.get("roofline",{}).get("frac"), r.get("e2e",{}).get("ms_per_step"))
    print(r.get("kernels_ms_per_step"))
except Exception as e:
    print("ERR",e); print(open(sys.argv[1].replace(".json",".err")).read()[-1200:])
PY
done
