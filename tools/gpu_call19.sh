#!/bin/bash
# GPU call 19: last-layer (9 -> 1) backward through the streaming kernel.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tp_fwd_bwd_explicit" > gpurun_out/r2t_pytest.log 2>&1; tail -2 gpurun_out/r2t_pytest.log
timeout 100 python tools/time_tp.py 2>&1 | head -8 > gpurun_out/r2t_time_tp.txt; cat gpurun_out/r2t_time_tp.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2t_bench_c2.json 2> gpurun_out/r2t_bench_c2.err
python - <<'PY'
import json
try:
    r=json.loads(open("gpurun_out/r2t_bench_c2.json").read().strip().splitlines()[-1])
    print(r["ms_per_step"], r["value"], r["e2e"]["ms_per_step"], r["parity_check"]); print(r["roofline"]); print(r["kernels_ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2t_bench_c2.err").read()[-1500:])
PY
