#!/bin/bash
# GPU call 18: quick stand-down, unroll-1 default of the three-warp backward, radial tile variant removed.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_zy_gpu_kernel_spec.py -q -m gpu -x -k "implicit_v0 or ragged or radial" > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
timeout 100 python tools/time_tp.py 2>&1 | head -4 > gpurun_out/r2r_time_tp.txt; cat gpurun_out/r2r_time_tp.txt
timeout 300 python bench.py > gpurun_out/r2r_bench_c2.json 2> gpurun_out/r2r_bench_c2.err
python - <<'PY'
import json
try:
    r=json.loads(open("gpurun_out/r2r_bench_c2.json").read().strip().splitlines()[-1])
    print(r["ms_per_step"], r["value"], r["e2e"]["ms_per_step"], r["parity_check"]); print(r["roofline"]); print(r["kernels_ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2r_bench_c2.err").read()[-1500:])
PY
timeout 300 python tools/time_triton_ref.py > gpurun_out/r2r_time_triton_ref.txt 2>&1; tail -11 gpurun_out/r2r_time_triton_ref.txt
