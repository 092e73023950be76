#!/bin/bash
# GPU call 14: new GPU tests -- trainable operator (weight grads, double backward), ZBL kernel + model.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_zbl.py -q -m gpu -x -k "contract or zbl or pair_potential" > gpurun_out/r2n_pytest_new.log 2>&1
tail -25 gpurun_out/r2n_pytest_new.log
