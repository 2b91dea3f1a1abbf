#!/bin/bash
# 8 GPUs: the nc = 1024 bench (strong scaling point) with NVLink counters around it, then BASELINE configs[4] (nc = 2048, 4096^3 mesh)
set -x
mkdir -p gpurun_out
nvidia-smi nvlink -gt d -i 0 > gpurun_out/r02_nvlink_before.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
nvidia-smi nvlink -gt d -i 0 > gpurun_out/r02_nvlink_after.txt 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_8gpu.json"))
    print(d["value"], d["ms_per_step"], d["e2e"], {k: (v["launches"], round(v["ms"], 1)) for k, v in d["stages_rank0"].items() if v["launches"]}, d["pk_bins"][:3], d["x_checksum"], d["np_total_after"])
except Exception as e:
    print("bench failed", e)
PY
tail -n 3 gpurun_out/r02_bench_8gpu.err
FASTPM_B200_MIGRATE_FRAC=0.01 FASTPM_B200_ARENA_FRAC=0.92 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29723 scripts/run_config.py --nc 2048 --pm-nc-factor 2 --steps 20 --mode pm --alloc 1.10 > gpurun_out/r02_c4_8gpu.json 2> gpurun_out/r02_c4_8gpu.err
cat gpurun_out/r02_c4_8gpu.json; tail -n 5 gpurun_out/r02_c4_8gpu.err
