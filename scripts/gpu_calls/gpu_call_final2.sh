#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -n 6 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_1gpu.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: (v["launches"], round(v["ms"] / max(1, v["launches"]), 2)) for k, v in d["stages"].items() if v["launches"]}, d["clocks"], d["pk_bins"][:3], d["x_checksum"])
PY
