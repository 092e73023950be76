#!/bin/bash
# GPU call 9: three-warp layer-0 backward (tp_stream3): parity, ragged rows, timing against the two-warp kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "implicit_v0 or ragged" > gpurun_out/r2i_pytest_tp.log 2>&1
tail -6 gpurun_out/r2i_pytest_tp.log
timeout 300 python tools/time_tp.py > gpurun_out/r2i_time_tp.txt 2>&1
head -4 gpurun_out/r2i_time_tp.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err
python - <<'PY'
import json
try:
    r=json.loads(open("gpurun_out/r2i_bench_c2.json").read().strip().splitlines()[-1])
    print(r["ms_per_step"], r["parity_check"], r["roofline"]); print(r["kernels_ms_per_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2i_bench_c2.err").read()[-1500:])
PY
