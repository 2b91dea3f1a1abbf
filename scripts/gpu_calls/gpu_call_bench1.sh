#!/bin/bash
# one-GPU bench line of the final code (end-to-end arm with overlapped copies)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02t_bench_1gpu.json 2> gpurun_out/r02t_bench_1gpu.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02t_bench_1gpu.json") if x.startswith("{")][-1]
d = json.loads(l)
print(d["value"], d["ms_per_step"], d["e2e"], {k: (v["launches"], round(v["ms"] / max(1, v["launches"]), 2)) for k, v in d["stages"].items() if v["launches"]}, d["clocks"], d["pk_bins"][:3], d["x_checksum"], d["roofline"]["frac"])
PY
tail -n 3 gpurun_out/r02t_bench_1gpu.err
