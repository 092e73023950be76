#!/bin/bash
# GPU call 15 (2 GPUs): gY-tile build of the layer-0 backward (GPU 0), halo parity on 2 GPUs, 2-GPU bench lines, c4-like box split in two.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "implicit_v0 or ragged" > gpurun_out/r2o_pytest_tp.log 2>&1; tail -3 gpurun_out/r2o_pytest_tp.log
timeout 120 python tools/time_tp.py 2>&1 | head -4 > gpurun_out/r2o_time_tp.txt; cat gpurun_out/r2o_time_tp.txt
timeout 400 python -m pytest tests/test_gpu_halo.py -q -m gpu > gpurun_out/r2o_pytest_halo.log 2>&1
tail -12 gpurun_out/r2o_pytest_halo.log
for halo in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 --halo $halo > gpurun_out/r2o_bench_2gpu_$halo.json 2> gpurun_out/r2o_bench_2gpu_$halo.err
  tail -c 1800 gpurun_out/r2o_bench_2gpu_$halo.json; echo; grep -E "Error|error" gpurun_out/r2o_bench_2gpu_$halo.err | head -5
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config c4 --reps 44 --steps 10 --warmup 3 > gpurun_out/r2o_bench_2gpu_c4r44.json 2> gpurun_out/r2o_bench_2gpu_c4r44.err
tail -c 1800 gpurun_out/r2o_bench_2gpu_c4r44.json; echo; grep -E "Error|error" gpurun_out/r2o_bench_2gpu_c4r44.err | head -5
