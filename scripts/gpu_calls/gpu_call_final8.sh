#!/bin/bash
# round 2, final state: full GPU suite and smoke()
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02u_pytest_gpu.log 2>&1; tail -n 12 gpurun_out/r02u_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02u_smoke.log 2>&1; tail -n 3 gpurun_out/r02u_smoke.log
