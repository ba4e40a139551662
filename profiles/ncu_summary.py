"""Per-launch summary of an `ncu --set full` report: duration, DRAM bytes, L2 hit rate, tensor / issue activity.

    python profiles/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > profiles/rNN_x.summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "us", "duration"),
    ("dram__bytes_read.sum", "MB", "dram read"),
    ("dram__bytes_write.sum", "MB", "dram write"),
    ("lts__t_sector_hit_rate.pct", "%", "L2 hit"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "%", "tensor pipe (hmma subpipe) active"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "%", "tensor instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%", "issue active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "%", "warps active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "%", "SM throughput"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%", "DRAM throughput"),
    ("launch__registers_per_thread", "", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "", "dynamic smem / block"),
]


def to_unit(val, unit, want):
    v = float(val.replace(",", ""))
    scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    if want == "us":
        return v * scale.get(unit, 1.0)
    if want == "MB":
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
    return v


for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(data)} launches (ncu --set full --import-source on --clock-control none; cold caches, kernels serialised)")
    for r in data:
        name = r[col["Kernel Name"]]
        grid, block = r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]
        print(f"{name[:110]}  grid {grid} block {block}")
        for key, want, label in WANT:
            if key in col and r[col[key]] not in ("", "n/a"):
                v = to_unit(r[col[key]], units[col[key]], want)
                print(f"    {label:<42}{v:12.2f} {want}")
    print()
