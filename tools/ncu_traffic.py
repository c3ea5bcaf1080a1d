"""Join an `ncu --set full` capture of tcgen05 GEMM launches with the engine's launch log (SUTA_GEMM_LOG) and write the
per-launch DRAM traffic next to the algorithmic bytes:

  on the GPU box:   SUTA_GEMM_LOG=gpurun_out/gemm_log.tsv ncu --set full --clock-control none --import-source on \
                        -k regex:gemm -s <skip> -c <n> -o gpurun_out/gemm_full python bench.py --steps 1 --warmup 1 ...
  here:             python tools/ncu_traffic.py gpurun_out/gemm_full.ncu-rep gpurun_out/gemm_log.tsv <skip> <workload> [out.md]

Writes / updates profiles/r02_gemm_traffic.json (read by bench.py for `roofline.traffic`) and prints a markdown table.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TUNIT = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}


def main():
    rep, log, skip, workload = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    if rep.endswith(".csv"):        # `ncu --csv --page raw --log-file x.csv` written on the GPU box (reports are too big to ship)
        raw = "".join(ln for ln in open(rep) if ln.startswith('"'))
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name, table):
        return float(r[ix[name]].replace(",", "")) * table.get(units[ix[name]], 1.0)

    launches = [ln.rstrip("\n").split("\t") for ln in open(log)]
    out = ["| # | kernel | M | N | K | nz | us | DRAM read MB | DRAM write MB | algorithmic MB | traffic / algorithmic | tensor pipe % |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    tot_t = tot_a = tot_us = 0.0
    n = 0
    for j, r in enumerate(body):
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:40]
        if skip + j >= len(launches):
            break
        _i, M, N, K, nz, alg, _fl = launches[skip + j]
        rd, wr = val(r, "dram__bytes_read.sum", UNIT), val(r, "dram__bytes_write.sum", UNIT)
        us = val(r, "gpu__time_duration.sum", TUNIT)
        tens = r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]] if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in ix else ""
        alg = float(alg)
        out.append(f"| {skip + j} | {name} | {M} | {N} | {K} | {nz} | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {alg / 1e6:.1f} | "
                   f"{(rd + wr) / alg:.2f} | {tens} |")
        tot_t += rd + wr; tot_a += alg; tot_us += us; n += 1
    out.append("")
    out.append(f"{n} launches: DRAM traffic {tot_t / n / 1e6:.1f} MB per launch, algorithmic {tot_a / n / 1e6:.1f} MB per launch, "
               f"ratio {tot_t / tot_a:.2f}; {tot_us / n:.1f} us per launch under ncu (cold caches, serialised)")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 5:
        with open(sys.argv[5], "w") as f:
            f.write(text + "\n")
    p = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    d[workload] = {"dram_bytes_per_launch": tot_t / n, "algorithmic_bytes_per_launch": tot_a / n, "launches_captured": n,
                   "note": f"ncu dram-byte capture of {n} consecutive tcgen05 GEMM launches (launch #{skip}..) of the benched batch, "
                           f"`bench.py --workload {workload} --steps 1 --warmup 1`; see profiles/ for the per-launch table"}
    json.dump(d, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
