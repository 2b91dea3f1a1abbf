#!/bin/bash
# round 2, first GPU call: microbenchmarks, FFT A/B (packed f32x2 vs scalar), C1 parity, bench, ncu pass-count probe
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt
free -g >> gpurun_out/r02_gpu.txt; nproc >> gpurun_out/r02_gpu.txt
./scripts/ubench/atomics > gpurun_out/r02_ubench_atomics.txt 2>&1
for n in 2048 1024; do
  python scripts/fft_passes.py $n > gpurun_out/r02_fft_passes_packed_$n.txt 2>&1
  python scripts/fft_passes.py $n scratch_ab/libfastpm_b200_scalar.so > gpurun_out/r02_fft_passes_scalar_$n.txt 2>&1
done
timeout 900 python -m pytest tests/test_gpu_c1.py -x -q -s -k "c1_matches_reference" > gpurun_out/r02_c1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver.py -x -q > gpurun_out/r02_gpu_tests_a.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu_a.json 2> gpurun_out/r02_bench_1gpu_a.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_probe_nc512.csv python bench.py --nc 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_probe_nc512.out 2>&1
tail -5 gpurun_out/r02_c1.log gpurun_out/r02_gpu_tests_a.log
cat gpurun_out/r02_ubench_atomics.txt
tail -12 gpurun_out/r02_fft_passes_*_2048.txt
head -c 1500 gpurun_out/r02_bench_1gpu_a.json
