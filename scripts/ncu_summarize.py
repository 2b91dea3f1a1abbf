#!/usr/bin/env python
"""Condense an ncu report (.ncu-rep, read with `ncu -i ... --page raw --csv`) into the per-kernel table kept under profiles/.

    python scripts/ncu_summarize.py gpurun_out/x.ncu-rep > profiles/rNN_x.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "ms", 1e-6),
    ("dram__bytes_read.sum", "GB read", 1e-9),
    ("dram__bytes_write.sum", "GB written", 1e-9),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__shared_mem_per_block_dynamic", "dyn smem", 1),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts", 1),
    ("smsp__inst_executed.sum", "warp inst", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1),
]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# ncu summary of `%s`\n" % rep.split("/")[-1])
    print("Per launch, `ncu --set full --clock-control none`; times are cold-cache and serialised (compare shares, not absolutes).")
    print("Achieved GB/s = (dram bytes read + written) / duration.\n")
    cols = ["kernel", "grid", "block"] + [w[1] for w in WANT if w[0] in ix] + ["DRAM GB/s"]
    print("| " + " | ".join(cols) + " |")
    print("|" + "---|" * len(cols))
    for r in data:
        name = r[ix["Kernel Name"]]
        name = name.split("(")[0].replace("void ", "")
        line = [name[:48], r[ix["Grid Size"]].replace(" ", ""), r[ix["Block Size"]].replace(" ", "")]
        vals = {}
        for key, label, scale in WANT:
            if key not in ix:
                continue
            v = to_float(r[ix[key]])
            u = units[ix[key]]
            if v is not None:
                if key.startswith("gpu__time_duration"):
                    v = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
                elif key.startswith("dram__bytes"):
                    v = v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}.get(u, 1e-9)
                elif key.endswith("dynamic"):
                    v = v * {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(u, 1.0)
            vals[key] = v
            line.append("-" if v is None else ("%.3f" % v if abs(v) < 1000 else "%.4g" % v))
        t, br, bw = vals.get("gpu__time_duration.sum"), vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
        line.append("%.0f" % ((br + bw) / (t * 1e-3)) if t and br is not None and bw is not None else "-")
        print("| " + " | ".join(line) + " |")


if __name__ == "__main__":
    main()
