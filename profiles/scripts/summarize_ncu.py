#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) and a `gpu__time_duration` launch list (.csv) into the small text
files kept under profiles/.  Runs on the CPU box: `ncu -i` only reads the report.

  python profiles/scripts/summarize_ncu.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv profiles/r01_xyz
"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 throughput %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "l2 read sectors"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
]


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep, launches, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = []
    if rep != "-":
        hdr, units, data = ncu_raw(rep)
        ki = hdr.index("Kernel Name")
        lines.append(f"# ncu --set full --clock-control none: {rep}\n")
        for r in data:
            name = r[ki].split("(")[0].replace("<unnamed>::", "")
            lines.append(f"## {name}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
            for key, label in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    lines.append(f"  {label:26s} {r[i]:>16s} {units[i]}")
            if "dram__bytes_read.sum" in hdr:
                pass
            lines.append("")
        open(prefix + "_ncu_full.txt", "w").write("\n".join(lines) + "\n")
    if launches != "-":
        rows = list(csv.reader(open(launches)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr = rows[hi]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        gi = hdr.index("Grid Size")
        agg = collections.OrderedDict()
        for r in rows[hi + 1:]:
            if len(r) <= vi:
                continue
            name = r[ki].split("(")[0].replace("<unnamed>::", "")
            agg.setdefault(name, []).append((float(r[vi].replace(",", "")), r[gi]))
        tot = sum(sum(v for v, _ in a) for a in agg.values())
        out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none: {launches}",
               "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
               f"{'kernel':28s} {'launches':>8s} {'avg us':>10s} {'total us':>11s} {'share':>7s}  grids"]
        for name, a in agg.items():
            t = sum(v for v, _ in a)
            grids = sorted(set(g for _, g in a))
            out.append(f"{name:28s} {len(a):8d} {t / len(a) / 1e3:10.2f} {t / 1e3:11.1f} {t / tot:7.1%}  {' '.join(grids[:4])}")
        open(prefix + "_launches.txt", "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
