#!/bin/bash
# GPU call 11 (2 GPUs): halo parity over NCCL and over NVLink peer memory, c2 slabs bench with both transports, c4-like box split in two.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2k_smi.txt
timeout 400 python -m pytest tests/test_gpu_halo.py -q -m gpu -x > gpurun_out/r2k_pytest_halo.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "contract" > gpurun_out/r2k_pytest_op.log 2>&1; tail -5 gpurun_out/r2k_pytest_op.log
tail -15 gpurun_out/r2k_pytest_halo.log
for halo in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 --halo $halo > gpurun_out/r2k_bench_2gpu_$halo.json 2> gpurun_out/r2k_bench_2gpu_$halo.err
  tail -c 1800 gpurun_out/r2k_bench_2gpu_$halo.json; echo; tail -n 4 gpurun_out/r2k_bench_2gpu_$halo.err
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config c4 --reps 44 --steps 10 --warmup 3 > gpurun_out/r2k_bench_2gpu_c4r44.json 2> gpurun_out/r2k_bench_2gpu_c4r44.err
tail -c 1800 gpurun_out/r2k_bench_2gpu_c4r44.json; echo; tail -n 6 gpurun_out/r2k_bench_2gpu_c4r44.err
