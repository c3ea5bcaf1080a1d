"""Summarise raw `ncu --page raw --csv` exports (one per kernel, tools/ncu_full_kernels.sh) into a markdown table:
python tools/ncu_csv_summary.py <dir> <out.md>"""
import csv
import io
import os
import sys

COLS = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("dram__bytes_read.sum", "rd MB"),
        ("dram__bytes_write.sum", "wr MB"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}


def main():
    d, out = sys.argv[1], sys.argv[2]
    lines = ["| kernel | " + " | ".join(n for _c, n in COLS) + " | top warp stalls (cycles per issue) |", "|---|" + "---|" * (len(COLS) + 1)]
    for f in sorted(os.listdir(d)):
        if not f.endswith(".csv"):
            continue
        raw = "".join(ln for ln in open(os.path.join(d, f)) if ln.startswith('"'))
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units, body = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_warp_active.pct") and "not_issued" not in h]
        for r in body[:2]:
            name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:48]
            vals = []
            for c, n in COLS:
                if c not in ix:
                    vals.append("-")
                    continue
                v = float(r[ix[c]].replace(",", "") or 0) * (UNIT.get(units[ix[c]], 1.0) if n in ("us", "rd MB", "wr MB") else 1.0)
                vals.append(f"{v:.1f}" if n not in ("regs", "grid", "block") else f"{v:.0f}")
            st = sorted(((float(r[ix[h]].replace(",", "") or 0), h.split("issue_stalled_")[1].split("_per_warp")[0]) for h in stall),
                        reverse=True)[:3]
            lines.append(f"| {name} | " + " | ".join(vals) + " | " + ", ".join(f"{n} {v:.0f} %" for v, n in st) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
