#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lua_front.py -x -q -s -m gpu > gpurun_out/r02i_lua.log 2>&1; tail -n 12 gpurun_out/r02i_lua.log
timeout 900 python -m pytest tests/test_gpu_c1.py -x -q -s -k "large_mesh" > gpurun_out/r02i_n1024.log 2>&1; tail -n 4 gpurun_out/r02i_n1024.log
