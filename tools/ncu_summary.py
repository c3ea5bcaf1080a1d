"""Compact per-launch summary of an .ncu-rep (read on the CPU box):  python tools/ncu_summary.py rep.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "rdMB"),
    ("dram__bytes_write.sum", "wrMB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_uniform.sum", "uinst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "blk"),
    ("smsp__inst_executed.sum", "inst"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    tens = [h for h in hdr if "tensor" in h and "pct" in h]
    lines = []
    lines.append("| # | kernel | " + " | ".join(n for c, n in COLS if c in idx) + " |")
    lines.append("|---|---|" + "---|" * sum(1 for c, _ in COLS if c in idx))
    for k, r in enumerate(body):
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        vals = []
        for c, n in COLS:
            if c not in idx:
                continue
            v = r[idx[c]]
            u = units[idx[c]]
            try:
                f = float(v.replace(",", ""))
                if n in ("rdMB", "wrMB"):
                    f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
                if n == "us":
                    f = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
                v = f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}"
            except ValueError:
                pass
            vals.append(v)
        lines.append(f"| {k} | {name[:60]} | " + " | ".join(vals) + " |")
    lines.append("")
    lines.append("tensor-pipe metrics present: " + ", ".join(tens))
    out = "\n".join(lines)
    print(out)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write(f"# ncu --set full summary of {rep}\n\n" + out + "\n")


if __name__ == "__main__":
    main()
