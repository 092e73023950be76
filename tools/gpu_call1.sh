#!/bin/bash
# GPU call 1 (round 2): full GPU test suite, baseline bench lines for c2 / c3 / c5, fold-embed A/B, micro-benchmarks.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -q -m gpu --maxfail=25 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
ALLEGRO_B200_FOLD_EMBED=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_c2_fold.json 2> gpurun_out/r2a_bench_c2_fold.err
timeout 900 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
timeout 900 python bench.py --config c5 --dtype float64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err
timeout 300 tools/ubench/ubench > gpurun_out/r2a_ubench.txt 2>&1
tail -5 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_ubench.txt | tail -70
for f in gpurun_out/r2a_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    r=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:r[k] for k in ("value","ms_per_step","dtype") if k in r}, r.get("parity_check"), r.get("roofline",{}).get("frac"), r.get("e2e",{}).get("ms_per_step"))
    print(r.get("kernels_ms_per_step"))
except Exception as e:
    print("ERR",e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
