#!/bin/bash
# round 2: two-GPU validation of the host-collective, wide-window and shifted-IC changes + sanitizer runs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multiprocess.py tests/test_gpu_c1.py tests/test_zz_first_gpu_run.py -m gpu -q -s -k "two_gpu" > gpurun_out/r02l_mp_tests.log 2>&1; grep -h "MP_.*OK\|passed\|failed\|skipped" gpurun_out/r02l_mp_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_zz_first_gpu_run.py -m gpu -q -k "window_and_derivative or shifted or non_cic" > gpurun_out/r02l_new_cases.log 2>&1; tail -n 3 gpurun_out/r02l_new_cases.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02l_bench_2gpu.json 2> gpurun_out/r02l_bench_2gpu.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02l_bench_2gpu.json") if x.startswith("{")][-1]
d = json.loads(l)
st = d.get("stages", d.get("stages_rank0"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("host_collectives"), {k: (v["launches"], round(v["ms"], 1)) for k, v in st.items() if v["launches"]}, d["clocks"], d["pk_bins"][:3], d["x_checksum"])
PY
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "paint_matches_reference or readout_bit_exact or window_and_derivative or kick_drift or powerspectrum_matches" > gpurun_out/r02l_sanitizer_memcheck.log 2>&1; tail -n 6 gpurun_out/r02l_sanitizer_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "r2c_c2r_match_reference or powerspectrum_matches" > gpurun_out/r02l_sanitizer_racecheck.log 2>&1; tail -n 6 gpurun_out/r02l_sanitizer_racecheck.log
