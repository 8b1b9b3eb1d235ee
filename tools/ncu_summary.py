"""Key figures of an `ncu --page raw --csv` export, one block per kernel launch:  python tools/ncu_summary.py file.csv ..."""
import csv, sys

KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__cluster_size", "cluster"),
        ("launch__registers_per_thread", "regs/thread"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (SM active)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
        ("sm__cycles_active.avg", "SM active cycles (avg)")]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print(f"{name[:110]}")
        for k, label in KEYS:
            if k in col and r[col[k]] != "":
                print(f"    {label:34s} {r[col[k]]:>16s} {units[col[k]]}")
