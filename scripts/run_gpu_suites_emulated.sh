#!/bin/bash
# Runs the GPU test files on the CPU against the emulated build of the whole library (tests/emul/emul_lib/).
# Not part of the default CPU suite: about 5 minutes for tests/test_gpu_kernels.py (the three N >= 512 FFT tests need the TMA path
# and are skipped), about 25 minutes for tests/test_gpu_solver.py at its full sizes.
#   scripts/run_gpu_suites_emulated.sh [pytest arguments, default: both files]
set -e
cd "$(dirname "$0")/.."
python tests/emul/emul_lib/build.py
export FASTPM_B200_TEST_EMUL=1
if [ $# -eq 0 ]; then
    set -- tests/test_gpu_kernels.py tests/test_gpu_solver.py
fi
exec python -m pytest -q -m gpu -p no:cacheprovider \
    --deselect tests/test_gpu_kernels.py::test_tma_fft_matches_generic_and_oracle \
    --deselect tests/test_gpu_kernels.py::test_fft_2048_round_trip_and_generic_planes "$@"
