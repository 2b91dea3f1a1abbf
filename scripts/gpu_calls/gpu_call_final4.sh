#!/bin/bash
# round 2, final state on two GPUs: multi-process parity tests, then the 2-GPU bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multiprocess.py tests/test_gpu_c1.py tests/test_zz_first_gpu_run.py -m gpu -q -s -k "two_gpu" > gpurun_out/r02k_mp_tests.log 2>&1; grep -h "MP_.*OK\|passed\|failed\|skipped" gpurun_out/r02k_mp_tests.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02k_bench_2gpu.json 2> gpurun_out/r02k_bench_2gpu.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02k_bench_2gpu.json") if x.startswith("{")][-1]
d = json.loads(l)
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: (v["launches"], round(v["ms"], 1)) for k, v in d.get("stages", d.get("stages_rank0")).items() if v["launches"]}, d["clocks"], d["pk_bins"][:3], d["x_checksum"], d.get("nvlink_rank0"))
PY
