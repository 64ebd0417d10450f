#!/usr/bin/env python
"""profiles/ncu_traffic.json from the round's `ncu --set full` summaries (scripts/ncu_summary.py CSVs).

usage: python scripts/make_traffic.py profiles/round2_ncu_full_k_gemm_cb.csv:gemm_cb[:TILE_BYTES] ...

Per kernel kind: the captured launches with their DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum); for the Schur
GEMM also the algorithmic bytes of the captured launch (every 64 x 64 tile reads and writes its piece of C once: 2 * 8 * 4096
bytes per CTA, operands come from L2) so that bench.py can scale the measured traffic / algorithmic ratio to its own launch mix."""
import csv, json, os, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def col(head, name):
    for i, h in enumerate(head):
        if h.split("[")[0] == name:
            u = h.split("[")[1].rstrip("]") if "[" in h else ""
            return i, UNIT.get(u, 1.0)
    raise KeyError(name)


out = {"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full captures of scripts/one_step.py lap3d 128 "
               "(profiles/round2_final_ncu_full_*.csv), cold caches", "kernels": {}}
for spec in sys.argv[1:]:
    parts = spec.split(":")
    path, kind = parts[0], parts[1]
    rows = list(csv.reader(open(path)))
    head, data = rows[0], rows[1:]
    ig, _ = col(head, "grid"); idur, udur = col(head, "duration_us"); ir, ur = col(head, "dram_read"); iw, uw = col(head, "dram_write")
    launches = []
    for r in data:
        grid = int(float(r[ig].split(",")[0].strip("() ")))
        d = {"grid": grid, "duration_us": float(r[idur]) * udur, "dram_bytes": float(r[ir]) * ur + float(r[iw]) * uw}
        if kind == "gemm_cb":
            d["algorithmic_bytes"] = grid * 2.0 * 8.0 * 4096.0
            d["traffic_over_algorithmic"] = d["dram_bytes"] / d["algorithmic_bytes"]
        launches.append(d)
    e = {"capture": os.path.basename(path), "launches": launches,
         "dram_bytes_per_launch_mean": sum(l["dram_bytes"] for l in launches) / max(1, len(launches))}
    if kind == "gemm_cb":
        e["traffic_over_algorithmic"] = sum(l["dram_bytes"] for l in launches) / sum(l["algorithmic_bytes"] for l in launches)
    out["kernels"][kind] = e
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: v.get("traffic_over_algorithmic", v["dram_bytes_per_launch_mean"]) for k, v in out["kernels"].items()}))
