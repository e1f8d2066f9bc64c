"""Summarise an .ncu-rep: key raw metrics and the top warp-stall SASS lines.
Usage: python tools/ncu_summary.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("=" * 100)
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} = {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = None
for n, r in enumerate(rows):
    if "Source" in r and "Warp Stall Sampling (All Samples)" in r:
        h = n
        break
if h is not None:
    hdr = rows[h]
    si, st, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    data = []
    for n, r in enumerate(rows[h + 1:]):
        try:
            data.append((int(r[st]), r[si][:100], int(r[ie]), n))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    print(f"--- top stall lines (of {tot} samples)")
    for d in sorted(data, key=lambda x: -x[0])[:top]:
        print(f"{d[3]:5d} {d[0]:6d} {100 * d[0] / tot:5.1f}%  ex={d[2]:>9d}  {d[1]}")
