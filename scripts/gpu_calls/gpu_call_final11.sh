#!/bin/bash
# two GPUs, final code: the slab run against the reference fixture and BASELINE configs[1] against the one-GPU run (staged transposes)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multiprocess.py tests/test_gpu_c1.py -m gpu -q -s -k "two_gpus_slab or (two_gpus_match and staged)" > gpurun_out/r02y_mp_tests.log 2>&1; grep -h "MP_.*OK\|passed\|failed\|skipped" gpurun_out/r02y_mp_tests.log | cut -c1-300
