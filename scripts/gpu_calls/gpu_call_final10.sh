#!/bin/bash
# N = 1536 with the per-pass tile chunking; the large-mesh FFT test again
set -x
mkdir -p gpurun_out
timeout 200 python scripts/fft_passes.py 1536 > gpurun_out/r02x_passes_1536.txt 2>&1; sed -n 1,7p gpurun_out/r02x_passes_1536.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fft_2048_round_trip or tma_fft_matches" > gpurun_out/r02x_fft_tests.log 2>&1; tail -n 3 gpurun_out/r02x_fft_tests.log
