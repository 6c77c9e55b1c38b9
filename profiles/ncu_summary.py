"""Summarise an `ncu --set full` report: python profiles/ncu_summary.py <rep.ncu-rep> > out.md"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor insts"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %")]
cols = [(hdr.index(k), n) for k, n in want if k in hdr]
ki = hdr.index("Kernel Name")
print("| kernel | " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for i, n in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in data:
    name = r[ki].split("(")[0].replace("void ", "")[:60]
    print(f"| `{name}` | " + " | ".join(r[i] for i, _ in cols) + " |")
