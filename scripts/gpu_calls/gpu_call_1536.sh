#!/bin/bash
# N = 1536 through the register/TMA passes (radix-24 first stage): correctness on a B200, pass timings against the generic passes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fft_2048_round_trip" > gpurun_out/r02p_fft_big.log 2>&1; tail -n 5 gpurun_out/r02p_fft_big.log
timeout 300 python scripts/fft_passes.py 1536 > gpurun_out/r02p_passes_1536_tma.txt 2>&1; tail -n 8 gpurun_out/r02p_passes_1536_tma.txt
FASTPM_B200_FFT=generic timeout 300 python scripts/fft_passes.py 1536 > gpurun_out/r02p_passes_1536_generic.txt 2>&1; tail -n 8 gpurun_out/r02p_passes_1536_generic.txt
