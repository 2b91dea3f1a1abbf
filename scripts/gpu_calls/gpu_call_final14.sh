#!/bin/bash
mkdir -p gpurun_out
timeout 30 python -m pytest tests/first_gpu_run_cases.py -m gpu -q -x -k "q_and_rand" > gpurun_out/r02z_rand.log 2>&1; tail -n 2 gpurun_out/r02z_rand.log
