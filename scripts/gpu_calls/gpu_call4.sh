#!/bin/bash
# 2 GPUs: multi-GPU parity tests and the 2-GPU bench with the new timing classes
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_c1.py tests/test_multiprocess.py tests/test_zz_first_gpu_run.py -x -q -s -k "two_gpus or two_gpu" > gpurun_out/r02d_mp_tests.log 2>&1; grep -E "MP_|passed|failed|Error" gpurun_out/r02d_mp_tests.log | tail -n 12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02d_bench_2gpu.json 2> gpurun_out/r02d_bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02d_bench_2gpu.json"))
print(d["value"], d["ms_per_step"], d["e2e"], {k: (v["launches"], round(v["ms"], 1)) for k, v in d["stages_rank0"].items() if v["launches"]}, d["pk_bins"][:3], d["x_checksum"], d["np_total_after"])
PY
tail -n 5 gpurun_out/r02d_bench_2gpu.err
