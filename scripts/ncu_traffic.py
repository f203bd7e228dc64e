"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) -> profiles/ncu_traffic_<tag>.json.
usage: ncu_traffic.py <ncu.csv> <bench.jsonl of the same command, run without ncu> <out.json>"""
import csv, json, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    if len(r) < len(hdr) or "fsv_fill" not in r[ix["Kernel Name"]]:
        continue
    key = "%s (id %s, grid %s)" % (r[ix["Kernel Name"]].replace("void fsv::", "").split("(fsv::")[0], r[ix["ID"]], r[ix["Grid Size"]])
    d = launches.setdefault(key, {})
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
    d[r[ix["Metric Name"]]] = val * mult
b = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
cells = b["roofline"]["cells_per_launch"] * max(1, round(b["gpu_launches"] / b["steps"])) if "roofline" in b else None
tot = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in launches.values())
out = {"command": "ncu --replay-mode application --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fsv_fill python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity",
       "workload": b["config"]["workload"], "launches": launches, "dram_bytes_per_step": tot,
       "cells_per_step": sum(v[1] for v in b["per_rank"]["values"]),
       "note": "DRAM bytes of the fill launches of ONE step (final round-2 build); algorithmic bytes = 1 B of traceback per in-band cell"}
out["bytes_per_cell"] = tot / out["cells_per_step"]
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps({k: out[k] for k in ("dram_bytes_per_step", "cells_per_step", "bytes_per_cell")}), len(launches), "launches")
