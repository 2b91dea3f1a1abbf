#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/fft_passes.py 2048 > gpurun_out/r02f_passes_x0.txt 2>&1
python scripts/fft_passes.py 2048 scratch_ab/libfastpm_b200_x1.so > gpurun_out/r02f_passes_x1.txt 2>&1
python scripts/fft_passes.py 2048 scratch_ab/libfastpm_b200_x2.so > gpurun_out/r02f_passes_x2.txt 2>&1
tail -n 7 gpurun_out/r02f_passes_x0.txt gpurun_out/r02f_passes_x1.txt gpurun_out/r02f_passes_x2.txt
