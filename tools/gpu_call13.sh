#!/bin/bash
# GPU call 13: tp_stream3 software-pipelined edge loop; compute-only knock-out.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "implicit_v0 or ragged" > gpurun_out/r2m_pytest_tp.log 2>&1
tail -4 gpurun_out/r2m_pytest_tp.log
timeout 300 python tools/time_tp3.py > gpurun_out/r2m_time_tp3.txt 2>&1
cat gpurun_out/r2m_time_tp3.txt
timeout 300 python tools/time_tp.py 2>&1 | head -3 > gpurun_out/r2m_time_tp.txt; cat gpurun_out/r2m_time_tp.txt
