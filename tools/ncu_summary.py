"""Markdown summary of an `ncu --set full` report:  ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]


def g(r, name, scale=1.0, fmt="{:.1f}"):
    if name not in hdr:
        return "-"
    try:
        return fmt.format(float(r[hdr.index(name)].replace(",", "")) * scale)
    except ValueError:
        return r[hdr.index(name)]


def unit(name):
    return units[hdr.index(name)] if name in hdr else ""


print("| # | kernel | grid x block | time (us) | tensor pipe active % | DRAM read (MB) | DRAM write (MB) | DRAM % of peak | L2 hit % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|")
for i, r in enumerate(data):
    name = r[hdr.index("Kernel Name")].replace("void ", "").replace("pn::", "")[:60]
    tu = unit("gpu__time_duration.sum")
    ts = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(tu, 1.0)

    def mb(n):
        u = unit(n)
        s = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        return g(r, n, s)
    grid = r[hdr.index("Grid Size")] if "Grid Size" in hdr else "?"
    blk = r[hdr.index("Block Size")] if "Block Size" in hdr else "?"
    print(f"| {i} | `{name}` | {grid} x {blk} | {g(r, 'gpu__time_duration.sum', ts)} | "
          f"{g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | {mb('dram__bytes_read.sum')} | "
          f"{mb('dram__bytes_write.sum')} | {g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | "
          f"{g(r, 'lts__t_sector_hit_rate.pct')} | {g(r, 'launch__registers_per_thread', 1.0, '{:.0f}')} |")
