#!/bin/bash
# round 2, final state: full GPU suite, default bench, reference arm
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02r_pytest_gpu.log 2>&1; tail -n 12 gpurun_out/r02r_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02r_bench_1gpu.json 2> gpurun_out/r02r_bench_1gpu.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02r_bench_1gpu.json") if x.startswith("{")][-1]
d = json.loads(l)
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: (v["launches"], round(v["ms"] / max(1, v["launches"]), 2)) for k, v in d["stages"].items() if v["launches"]}, d["clocks"], d["pk_bins"][:3], d["x_checksum"], d["roofline"])
PY
timeout 600 python bench.py --impl reference > gpurun_out/r02r_bench_ref.json 2> gpurun_out/r02r_bench_ref.err; head -c 700 gpurun_out/r02r_bench_ref.json; echo
