#!/bin/bash
# do kz-adjacent tiles taken back to back by one CTA (the two 64-byte halves of a line written microseconds apart) help the transposing passes?
set -x
mkdir -p gpurun_out
for c in 1 2 4; do
  FASTPM_B200_TMA_CHUNK=$c timeout 200 python scripts/fft_passes.py 1536 > gpurun_out/r02w_passes_1536_chunk$c.txt 2>&1; echo "1536 chunk $c"; sed -n 1,5p gpurun_out/r02w_passes_1536_chunk$c.txt
  FASTPM_B200_TMA_CHUNK=$c timeout 200 python scripts/fft_passes.py 2048 > gpurun_out/r02w_passes_2048_chunk$c.txt 2>&1; echo "2048 chunk $c"; sed -n 1,5p gpurun_out/r02w_passes_2048_chunk$c.txt
done
