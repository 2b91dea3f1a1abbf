#!/bin/bash
# super-block traversal of paint / readout at nc = 1024: last-step kernel times for several block shapes
set -x
mkdir -p gpurun_out
for sb in 1x1 8x4 16x8 4x2 32x16; do
  FASTPM_B200_SUPERBLOCK=$sb timeout 300 python bench.py --steps 4 --warmup 1 --no-cpu-baseline > gpurun_out/r02g_sb_$sb.json 2> gpurun_out/r02g_sb_$sb.err
done
python - <<'PY'
import json
for sb in ["1x1", "8x4", "16x8", "4x2", "32x16"]:
    try:
        d = json.load(open("gpurun_out/r02g_sb_%s.json" % sb))
        ls = d["last_step_launches"]
        print(sb, round(d["ms_per_step"], 1), "paint avg", round(d["stages"]["paint"]["ms"] / d["stages"]["paint"]["launches"], 2), "readout avg", round(d["stages"]["readout"]["ms"] / d["stages"]["readout"]["launches"], 2),
              "last step:", [x for x in ls if x[0] in ("paint", "readout")], d["x_checksum"][0])
    except Exception as e:
        print(sb, "failed", e)
PY
