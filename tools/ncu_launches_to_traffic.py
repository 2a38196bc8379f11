"""ncu launch list (csv, --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum) of a bench.py run ->
per-kernel-family DRAM bytes and time of ONE timed pass (profiles/r2_dram_traffic.json).
    python tools/ncu_launches_to_traffic.py gpurun_out/r2_ncu_launches_bench.csv <chunk> [pass_index] > profiles/r2_dram_traffic.json
A pass = the launches from one preprocess kernel up to (not including) the next one."""
import collections
import csv
import json
import re
import sys

path, chunk = sys.argv[1], int(sys.argv[2])
want_pass = int(sys.argv[3]) if len(sys.argv) > 3 else -1
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
launches = collections.OrderedDict()
for r in rows:
    d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"].startswith("gpu__time_duration"):
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else v * ({"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1))
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d["rd" if "read" in r["Metric Name"] else "wr"] = v * mult


def family(name):
    m = re.search(r"(\w+_kernel(?:<\d+>)?)", name.replace("ub::", "").replace("(int)", ""))
    return m.group(1) if m else name[:40]


ids = list(launches)
starts = [i for i in ids if "preprocess" in launches[i]["name"]]
passes = []
for a, b in zip(starts, starts[1:] + [ids[-1] + 1]):
    p = []
    for i in ids:
        if a <= i < b:
            if "pack_" in launches[i]["name"] or "at::native" in launches[i]["name"]:
                break            # the next plan's weight packing / a torch fill: not part of the pass
            p.append(launches[i])
    if len(p) >= 20:
        passes.append(p)
if not passes:
    sys.exit("no complete pass found")
p = passes[want_pass]
fams = collections.OrderedDict()
tot_us = sum(l.get("us", 0.0) for l in p)
for l in p:
    f_ = fams.setdefault(family(l["name"]), {"launches": 0, "dram_read_bytes": 0, "dram_write_bytes": 0, "ncu_us": 0.0})
    f_["launches"] += 1
    f_["dram_read_bytes"] += int(l.get("rd", 0))
    f_["dram_write_bytes"] += int(l.get("wr", 0))
    f_["ncu_us"] += l.get("us", 0.0)
for f_ in fams.values():
    f_["share_of_pass"] = f_["ncu_us"] / tot_us
out = {"source": f"{path}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                 "python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --no-fp32 --no-cfg5 (B200, round 2)",
       "unit": f"bytes per {chunk}-frame pass at 224x224 (sum over the launches of the family); ncu_us is cold-cache serialised time (compare shares)",
       "chunk": chunk, "hw": [224, 224], "passes_found": len(passes), "pass_us_under_ncu": tot_us,
       "pass_dram_read_bytes": sum(f_["dram_read_bytes"] for f_ in fams.values()),
       "pass_dram_write_bytes": sum(f_["dram_write_bytes"] for f_ in fams.values()), "families": fams,
       "launch_order": [family(l["name"]) for l in p]}
print(json.dumps(out, indent=1))
