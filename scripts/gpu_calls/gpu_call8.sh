#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lua_front.py -x -q -s -m gpu > gpurun_out/r02h_lua.log 2>&1; tail -n 15 gpurun_out/r02h_lua.log
mkdir -p /tmp/c0 && cp tests/golden/powerspec.txt /tmp/c0/ && (cd /tmp/c0 && timeout 300 /root/repo/fastpm_b200/lua_front/_build/fastpm_b200_run /root/repo/tests/lua/c0_standard.lua cola > /root/repo/gpurun_out/r02h_c0_cola.log 2>&1; echo "c0 rc=$?"; ls c0-cola | head)
grep -E "Force Calculation|D\^2|written|dx1|dx2" gpurun_out/r02h_c0_cola.log | head -n 20
timeout 900 python -m pytest tests/test_gpu_c1.py -x -q -s -k "large_mesh" > gpurun_out/r02h_n1024.log 2>&1; tail -n 4 gpurun_out/r02h_n1024.log
timeout 300 python scripts/run_config.py --nc 256 --pm-nc-factor "0:1,0.5:3" --steps 8 --mode fastpm > gpurun_out/r02h_vpm_small.json 2> gpurun_out/r02h_vpm_small.err; cat gpurun_out/r02h_vpm_small.json; tail -n 3 gpurun_out/r02h_vpm_small.err
