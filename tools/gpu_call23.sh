#!/bin/bash
# GPU call 23: radial_pq_bwd with float4 PQ loads -- kernel tests + default bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zy_gpu_kernel_spec.py tests/test_gpu_model.py -q -m gpu -x > gpurun_out/r2w_pytest.log 2>&1; tail -3 gpurun_out/r2w_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2w_bench_c2.json 2> gpurun_out/r2w_bench_c2.err
python - <<'PY'
import json
try:
    r=json.loads(open("gpurun_out/r2w_bench_c2.json").read().strip().splitlines()[-1])
    print(r["ms_per_step"], r["value"], r["e2e"]["ms_per_step"], r["parity_check"]); print(r["kernels_ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2w_bench_c2.err").read()[-1500:])
PY
