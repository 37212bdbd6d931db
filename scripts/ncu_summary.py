"""Summarise `ncu --page raw --csv` exports: one line per launch with the metrics DESIGN.md cites."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]][:60]
        out = [f"{name:60s}"]
        for k, short in KEYS:
            if k in col:
                v, u = r[col[k]], units[col[k]]
                try:
                    f = float(v.replace(",", ""))
                    if short == "dur":
                        f = f / 1e6 if u in ("ns", "nsecond") else (f / 1e3 if u.startswith("us") else f)
                        out.append(f"{short}={f:.3f}ms")
                    elif short.startswith("dram_"):
                        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                        out.append(f"{short}={f * scale / 1e9:.3f}GB")
                    elif short in ("inst",):
                        out.append(f"{short}={f:.3g}")
                    else:
                        out.append(f"{short}={f:.4g}")
                except ValueError:
                    out.append(f"{short}={v}")
        print(" ".join(out))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("##", p)
        main(p)
