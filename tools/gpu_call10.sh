#!/bin/bash
# GPU call 10: what bounds the layer-0 backward?  stage knock-outs + ncu source-level capture of tp_bwd3_kernel.
mkdir -p gpurun_out
timeout 300 python tools/time_tp3.py > gpurun_out/r2j_time_tp3.txt 2>&1
cat gpurun_out/r2j_time_tp3.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tp_bwd3_kernel" -c 2 -o gpurun_out/r2j_bwd3 python tools/prof_one.py > gpurun_out/r2j_ncu.log 2>&1
tail -3 gpurun_out/r2j_ncu.log; ls -la gpurun_out/
