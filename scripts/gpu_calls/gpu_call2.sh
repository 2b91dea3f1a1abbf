#!/bin/bash
# round 2, second GPU call: kernels + solver tests with the tile kernels / P(k) rows, C1 parity, A/B of the tiles at nc=1024
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver.py -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02b_gpu_tests.log
timeout 900 python -m pytest tests/test_gpu_c1.py -x -q -s -k "c1_matches_reference" > gpurun_out/r02b_c1.log 2>&1; tail -n 4 gpurun_out/r02b_c1.log
python scripts/fft_passes.py 2048 > gpurun_out/r02b_fft_passes_2048.txt 2>&1; tail -n 5 gpurun_out/r02b_fft_passes_2048.txt
FASTPM_B200_TILE_STATS=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_tiles.json 2> gpurun_out/r02b_bench_tiles.err
FASTPM_B200_TILES=0 FASTPM_B200_PK=generic timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_notiles.json 2> gpurun_out/r02b_bench_notiles.err
python - <<'PY'
import json
for f in ["gpurun_out/r02b_bench_tiles.json", "gpurun_out/r02b_bench_notiles.json"]:
    try:
        d = json.load(open(f))
        print(f, d["ms_per_step"], {k: (v["launches"], round(v["ms"] / max(1, v["launches"]), 2)) for k, v in d["stages"].items() if v["launches"]}, d.get("tile_stats"), d["pk_bins"][:3], d["x_checksum"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -n 5 gpurun_out/r02b_bench_tiles.err
