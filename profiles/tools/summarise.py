"""Turn the `ncu --set full` reports of tools/profile_all.sh into one markdown table.
    python profiles/tools/summarise.py gpurun_out/prof_*.ncu-rep > profiles/r1/kernels.md
Numbers taken under ncu are cold-cache, serialised replays: they explain the bench values, they are not bench values."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
KEYS = {
    "t_us": "gpu__time_duration.sum", "regs": "launch__registers_per_thread", "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "fp64_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "inst": "smsp__inst_executed.sum",
    "lld": "sass__inst_executed_local_loads", "name": "Kernel Name", "grid": "Grid Size", "block": "Block Size",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2]
    return {h: (data[i], units[i]) for i, h in enumerate(hdr)}


def num(m, key):
    v, u = m.get(KEYS[key], ("nan", ""))
    try:
        return float(v.replace(",", "")) * UNIT.get(u, 1.0)
    except ValueError:
        return float("nan")


def main(reps):
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = peaks.get("hbm_gbs", 6554.2)
    print("| capture | kernel | grid x block | regs | time (us) | DRAM R+W (MB) | DRAM GB/s (%% of measured %.0f) | fp64 pipe %% | fma pipe %% | issue %% | warps active %% | warp instr | local loads |" % hbm)
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for rep in reps:
        m = raw(rep)
        t, rd, wr = num(m, "t_us"), num(m, "rd"), num(m, "wr")
        gbs = (rd + wr) / (t * 1e-6) / 1e9
        name = m.get(KEYS["name"], ("?", ""))[0].split("(")[0][:48]
        print("| %s | `%s` | %s x %s | %d | %.1f | %.1f | %.0f (%.0f %%) | %.1f | %.1f | %.1f | %.1f | %.3g | %.3g |" % (
            os.path.basename(rep).replace("prof_", "").replace(".ncu-rep", ""), name, m.get(KEYS["grid"], ("?",))[0],
            m.get(KEYS["block"], ("?",))[0], num(m, "regs"), t, (rd + wr) / 1e6, gbs, 100 * gbs / hbm, num(m, "fp64_pct"),
            num(m, "fma_pct"), num(m, "issue_pct"), num(m, "warps_pct"), num(m, "inst"), num(m, "lld")))


if __name__ == "__main__":
    main(sys.argv[1:])
