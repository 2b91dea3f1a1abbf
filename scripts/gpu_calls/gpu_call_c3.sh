#!/bin/bash
# 4 GPUs: BASELINE configs[3] -- nc = 512, force mesh 512^3 -> 1536^3 at a = 0.5 (vpm.c), 40 steps
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29731 scripts/run_config.py --nc 512 --pm-nc-factor "0:1,0.5:3" --steps 40 --mode fastpm > gpurun_out/r02_c3_4gpu.json 2> gpurun_out/r02_c3_4gpu.err
cat gpurun_out/r02_c3_4gpu.json; tail -n 4 gpurun_out/r02_c3_4gpu.err
