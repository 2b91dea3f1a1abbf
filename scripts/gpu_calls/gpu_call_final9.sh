#!/bin/bash
# ncu --set full of the N = 1536 passes (radix-24 first stage): one forward and one inverse transform with the force kernel
set -x
mkdir -p gpurun_out
cat > /tmp/one_1536.py <<'PY'
import sys
sys.path.insert(0, "/root/repo")
from fastpm_b200 import _lib, device
n = 1536
m = device.Mesh(n, float(n) / 2)
a, b = device.DeviceBuffer(m.alloc_floats * 4), device.DeviceBuffer(m.alloc_floats * 4)
_lib.check(m.lib.fpm_fill_whitenoise(m.h, a.ptr, 1), "noise")
m.r2c(a, b)
m.c2r(b, a, m.transfer_for_kernel("1_4", 0, 0))
_lib.check(m.lib.fpm_sync(), "sync")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fft_tma_kernel|fft_zrow_kernel' -o gpurun_out/r02v_fft_n1536 python /tmp/one_1536.py > gpurun_out/r02v_ncu_1536.log 2>&1; tail -n 3 gpurun_out/r02v_ncu_1536.log
ls -la gpurun_out/r02v_fft_n1536.ncu-rep
