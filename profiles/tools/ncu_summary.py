"""Summarise an .ncu-rep: key metrics + stall reasons (reads the raw page via the ncu CLI)."""
import csv, subprocess, sys
def summary(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
                "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
                "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
                "l1tex__t_bytes.sum", "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active"]
        for k in keys:
            if k in d: res.append(f"{k:75s} {d[k][0]} {d[k][1]}")
        st = []
        for h, (v, u) in d.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try: st.append((float(v.replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError: pass
        st.sort(reverse=True)
        res.append("stalls per issue: " + ", ".join(f"{h}={v:.2f}" for v, h in st[:9]))
    return "\n".join(res)
if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p); print(summary(p))
