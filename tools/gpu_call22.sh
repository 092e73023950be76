#!/bin/bash
# GPU call 22: ncu --set full of the kernels added at the end of the round (9 -> 1 streaming backward, baked fp64 tensor products).
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"tp_baked64_kernel|tp_stream_kernel" -c 8 -o gpurun_out/r2v_prof python tools/prof_two.py > gpurun_out/r2v_ncu.log 2>&1
tail -3 gpurun_out/r2v_ncu.log; ls -la gpurun_out | tail -4
