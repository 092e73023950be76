#!/bin/bash
# GPU call 16 (8 GPUs): the driver's own N=8 command (c2 weak scaling + the 1M-atom c4 box split into 8 slabs under "c4"),
# halo parity on 2 of the GPUs first.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r2p_smi.txt
timeout 200 python -m pytest tests/test_gpu_halo.py -q -m gpu > gpurun_out/r2p_pytest_halo.log 2>&1
tail -4 gpurun_out/r2p_pytest_halo.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2p_bench_8gpu.json 2> gpurun_out/r2p_bench_8gpu.err
tail -c 3000 gpurun_out/r2p_bench_8gpu.json; echo; grep -E "Error|error|p2p halo" gpurun_out/r2p_bench_8gpu.err | head -8
