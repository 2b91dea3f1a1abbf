#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_1gpu_b.json 2> gpurun_out/r02_bench_1gpu_b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_1gpu_b.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: (v["launches"], round(v["ms"] / max(1, v["launches"]), 2)) for k, v in d["stages"].items() if v["launches"]}, d["clocks"], d["pk_bins"][:3], d["x_checksum"])
print(d["last_step_launches"])
PY
timeout 600 python -m pytest tests/test_gpu_solver.py tests/test_gpu_kernels.py -q -x > gpurun_out/r02_q_tests.log 2>&1; tail -n 3 gpurun_out/r02_q_tests.log
