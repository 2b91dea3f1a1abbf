#!/bin/bash
# 8 GPUs, final code of round 2: the nc = 1024 bench (strong-scaling point)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02o_bench_8gpu.json 2> gpurun_out/r02o_bench_8gpu.err
python - <<'PY'
import json
try:
    l = [x for x in open("gpurun_out/r02o_bench_8gpu.json") if x.startswith("{")][-1]
    d = json.loads(l)
    print(d["value"], d["ms_per_step"], d["e2e"], d.get("host_collectives"), {k: (v["launches"], round(v["ms"], 1)) for k, v in d["stages_rank0"].items() if v["launches"]}, d["pk_bins"][:3], d["x_checksum"], d["np_total_after"], d["clocks"])
except Exception as e:
    print("bench failed", e)
PY
tail -n 3 gpurun_out/r02o_bench_8gpu.err
