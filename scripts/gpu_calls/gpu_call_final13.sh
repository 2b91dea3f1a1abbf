#!/bin/bash
# the command line after its last changes (passive transition log, -n parsing): golden log of the reference on a B200
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_lua_front.py -m gpu -q -x > gpurun_out/r02z_lua.log 2>&1; tail -n 3 gpurun_out/r02z_lua.log
