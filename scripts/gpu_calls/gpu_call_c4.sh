#!/bin/bash
# 8 GPUs: BASELINE configs[4] -- nc = 2048, 4096^3 mesh, 2LPT initial conditions + 20 plain-PM steps (HBM-capacity run)
set -x
mkdir -p gpurun_out
FASTPM_B200_MIGRATE_FRAC=0.01 FASTPM_B200_ARENA_FRAC=0.92 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29723 scripts/run_config.py --nc 2048 --pm-nc-factor 2 --steps 20 --mode pm --alloc 1.10 > gpurun_out/r02_c4_8gpu.json 2> gpurun_out/r02_c4_8gpu.err
cat gpurun_out/r02_c4_8gpu.json; tail -n 6 gpurun_out/r02_c4_8gpu.err
