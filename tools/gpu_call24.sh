#!/bin/bash
# GPU call 24: last confirmation -- operator tests, reference-generated golden vectors, default bench line incl. the CPU baseline.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_zx_gpu_reference_golden.py tests/test_zbl.py -q -m gpu -x -k "contract or golden or zbl or pair" > gpurun_out/r2x_pytest.log 2>&1; tail -3 gpurun_out/r2x_pytest.log
timeout 300 python bench.py > gpurun_out/r2x_bench_c2.json 2> gpurun_out/r2x_bench_c2.err
python - <<'PY'
import json
try:
    r=json.loads(open("gpurun_out/r2x_bench_c2.json").read().strip().splitlines()[-1])
    print(r["ms_per_step"], r["value"], r["e2e"]["ms_per_step"], r["parity_check"], r["cpu_baseline"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2x_bench_c2.err").read()[-1500:])
PY
