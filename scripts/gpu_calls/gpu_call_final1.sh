#!/bin/bash
# round 2: full GPU suite, default bench, reference arm, launch list and ncu captures for profiles/
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -n 5 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; head -c 600 gpurun_out/r02_bench_1gpu.json; echo
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; head -c 1200 gpurun_out/r02_bench_ref.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_nc1024.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cic_paint_kernel|cic_readout_kernel|powerspectrum_rows|fused_update|fft_tma_kernel|fft_zrow_kernel' --launch-skip 40 --launch-count 16 -o gpurun_out/r02_force_nc512 python scripts/profile_step.py 512 2 > gpurun_out/r02_ncu_nc512.log 2>&1; tail -n 2 gpurun_out/r02_ncu_nc512.log
timeout 420 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'cic_paint_kernel|cic_readout_kernel|powerspectrum_rows|fused_update' --launch-skip 6 --launch-count 7 --csv --log-file gpurun_out/r02_traffic_particles_nc1024.csv python scripts/profile_step.py 1024 2 > gpurun_out/r02_traffic_particles.out 2>&1; tail -n 3 gpurun_out/r02_traffic_particles.out; grep -c "dram__bytes" gpurun_out/r02_traffic_particles_nc1024.csv
