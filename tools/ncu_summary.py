"""ncu raw-page CSV (`ncu -i x.ncu-rep --page raw --csv`) -> one line per captured launch with the metrics the roofline
discussion uses.  Usage: python tools/ncu_summary.py raw.csv > summary.txt"""
import csv, sys

WANT = [("gpu__time_duration.sum", "time_us", 1e-3), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1), ("launch__registers_per_thread", "regs", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%", 1), ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy%(elapsed)", 1), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/instr", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%", 1),
        ("dram__bytes_read.sum", "dram_rd_MB", 1e-6), ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts", 1)]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    names = [w for w in WANT if w[0] in col]
    print("%-28s" % "kernel" + "".join("%18s" % w[1] for w in names))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        out = "%-28s" % r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("h2v::", "")[:27]
        for key, _label, scale in names:
            v, u = r[col[key]].replace(",", ""), units[col[key]]
            try:
                x = float(v) * scale
                if key == "gpu__time_duration.sum":
                    x = float(v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                if key.startswith("dram__bytes"):
                    x = float(v) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                out += "%18.2f" % x
            except ValueError:
                out += "%18s" % v[:16]
        print(out)


if __name__ == "__main__":
    main()
