"""One line per kernel launch from an ncu report (raw page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
# an .ncu-rep, or the `ncu -i rep --page raw --csv` dump of one (made on the GPU box when the report is too large to bring back)
txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(r, name):
    try: return r[hdr.index(name)]
    except ValueError: return "-"
print(f"{'kernel':34s} {'ms':>8s} {'grid':>6s} {'blk':>5s} {'regs':>4s} {'occ%':>5s} {'issue%':>6s} {'lanes':>5s} {'dramR MB':>9s} {'dramW MB':>9s} {'dram%':>6s} {'winst M':>8s} {'bankx M':>8s}")
for r in data:
    name = col(r, "Kernel Name").split("(")[0].replace("void ", "").replace("rspt::", "")[:34]
    def f(n, scale=1.0):
        try: return float(col(r, n).replace(",", "")) * scale
        except ValueError: return float("nan")
    print(f"{name:34s} {f('gpu__time_duration.sum'):8.3f} {col(r,'launch__grid_size'):>6s} {col(r,'launch__block_size'):>5s} {col(r,'launch__registers_per_thread'):>4s} "
          f"{f('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {f('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f} "
          f"{f('dram__bytes_read.sum'):9.1f} {f('dram__bytes_write.sum'):9.1f} {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {f('smsp__inst_executed.sum',1e-6):8.1f} {f('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',1e-6):8.2f}")
print("units:", {n: units[hdr.index(n)] for n in ('gpu__time_duration.sum','dram__bytes_read.sum') if n in hdr})
