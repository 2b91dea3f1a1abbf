#!/bin/bash
# experiment: one pass over the particles for the three force components on 2 GPUs (FASTPM_B200_FUSED_READOUT=1, no pipelining)
set -x
mkdir -p gpurun_out
FASTPM_B200_FUSED_READOUT=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 > gpurun_out/r02m_bench_2gpu_fused.json 2> gpurun_out/r02m_bench_2gpu_fused.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02m_bench_2gpu_fused.json") if x.startswith("{")][-1]
d = json.loads(l)
st = d.get("stages", d.get("stages_rank0"))
print(d["value"], d["ms_per_step"], {k: (v["launches"], round(v["ms"], 1)) for k, v in st.items() if v["launches"]}, d["pk_bins"][:3])
PY
tail -n 3 gpurun_out/r02m_bench_2gpu_fused.err
