#!/bin/bash
# the three GPU cases added after the last full run: passive handlers, q / rand columns, tidal transfers
mkdir -p gpurun_out
timeout 55 python -m pytest tests/first_gpu_run_cases.py tests/test_gpu_kernels.py -m gpu -q -x -k "passive or q_and_rand or gravity_kernel_bit_exact" > gpurun_out/r02z_new_cases.log 2>&1; tail -n 4 gpurun_out/r02z_new_cases.log
