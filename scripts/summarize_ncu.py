"""Summarises an ncu report (read here, no GPU needed) into a text file for profiles/.

    python scripts/summarize_ncu.py gpurun_out/<name>.ncu-rep profiles/<name>_summary.txt [workload-key]

Writes per-kernel duration, DRAM bytes, throughput, occupancy, issue utilisation and the warp-stall
distribution; with a workload key (e.g. cfg4/k128/f32) -- the report then has to hold exactly the launches of
ONE solve -- also records their summed DRAM traffic in profiles/ncu_traffic.json, which bench.py reports as
roofline.traffic.
"""
import csv
import json
import os
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic"]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(value) * scale


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = ["ncu --set full --clock-control none, report %s" % os.path.basename(rep), ""]
    name_i = hdr.index("Kernel Name")
    traffic = []
    for r in data:
        lines.append("kernel %s" % r[name_i])
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                lines.append("  %-60s %s %s" % (m, r[i], units[i]))
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                lines.append("  stall %-40s %.2f warps per issue" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(r[i])))
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic.append(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]))
        lines.append("  dram traffic (read + write)                                  %.3f GB" % (traffic[-1] / 1e9))
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines))
    if key and traffic:
        path = os.path.join(os.path.dirname(os.path.abspath(out)), "ncu_traffic.json")
        table = json.load(open(path)) if os.path.exists(path) else {}
        gi = hdr.index("launch__grid_size")
        names = [r[name_i].split("(")[0] for r in data]
        main = [r for r in data if r[name_i].startswith(("chb_sweep", "chb_front"))] or data
        table[key] = {"bytes_per_step": sum(traffic), "launches": len(traffic), "grid": int(float(main[0][gi])),
                      "kernels": " ".join(names), "source": "profiles/" + os.path.basename(out)}
        json.dump(table, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
