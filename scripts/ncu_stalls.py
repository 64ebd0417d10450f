#!/usr/bin/env python
"""Per-launch summary + top stall reasons from an ncu --set full report."""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = rows[0]
for row in rows[2:]:
    g = lambda n: row[h.index(n)] if n in h else "?"
    print("==", g("Kernel Name").split("::")[-1][:40], "grid", g("launch__grid_size"), "dur_us", g("gpu__time_duration.sum"),
          "dram_rd", g("dram__bytes_read.sum"), "wr", g("dram__bytes_write.sum"), "dram%", g("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
          "sm%", g("sm__throughput.avg.pct_of_peak_sustained_elapsed"), "occ%", g("sm__warps_active.avg.pct_of_peak_sustained_active"),
          "inst", g("smsp__inst_executed.sum"), "ipc", g("sm__inst_executed.avg.per_cycle_active"), "l1hit%", g("l1tex__t_sector_hit_rate.pct"), "l2hit%", g("lts__t_sector_hit_rate.pct"))
    st = []
    for i, n in enumerate(h):
        if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued"):
            try: st.append((float(row[i]), n.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError: pass
    st.sort(reverse=True); tot = sum(v for v, _ in st) or 1
    print("   stalls:", ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in st[:6]))
