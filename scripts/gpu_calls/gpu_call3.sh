#!/bin/bash
set -x
mkdir -p gpurun_out
FASTPM_B200_TILES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cic_paint_tile|cic_readout_tile|powerspectrum_rows' --launch-skip 6 --launch-count 5 -o gpurun_out/r02c_tiles_nc512 python scripts/profile_step.py 512 2 > gpurun_out/r02c_ncu.log 2>&1; tail -n 3 gpurun_out/r02c_ncu.log
timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c_bench.json"))
print(d["ms_per_step"], {k: (v["launches"], round(v["ms"] / max(1, v["launches"]), 2)) for k, v in d["stages"].items() if v["launches"]}, d["pk_bins"][:3])
print(d["last_step_launches"])
PY
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "power or pk or paint or readout" > gpurun_out/r02c_tests.log 2>&1; tail -n 3 gpurun_out/r02c_tests.log
