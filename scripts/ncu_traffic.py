"""profiles/r02_traffic.json from an `ncu --set full` capture of the handler kernels (k_node_msgs + k_node_tasks) over
consecutive ticks around t = 1 600 ms of the GSF-131072 run: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and
duration per launch, and the per-tick average that bench.py reports as roofline.traffic (19 ordinary ticks + 1 doCycle tick
per period of 20 ms).  usage: ncu_traffic.py <raw.csv> (from `ncu -i x.ncu-rep --page raw --csv`)"""
import csv
import json
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
units = rows[1]


def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1,
             "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1}.get(u, 1)
    return v * scale


launches = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0]
    launches.append({"kernel": name, "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                     "seconds": val(r, "gpu__time_duration.sum")})
# pair msgs + tasks per tick
ticks = []
for i in range(0, len(launches) - 1, 2):
    a, b = launches[i], launches[i + 1]
    ticks.append({"dram_bytes": a["dram_bytes"] + b["dram_bytes"], "seconds": a["seconds"] + b["seconds"]})
big = max(ticks, key=lambda t: t["dram_bytes"])
small = min(ticks, key=lambda t: t["dram_bytes"])
has_cycle = big["dram_bytes"] > 4 * small["dram_bytes"]  # a doCycle tick (every 20th) moves an order of magnitude more
ordinary = [t for t in ticks if not has_cycle or t["dram_bytes"] < 0.5 * big["dram_bytes"]]
mean_o = sum(t["dram_bytes"] for t in ordinary) / len(ordinary)
sec_o = sum(t["seconds"] for t in ordinary) / len(ordinary)
out = {"k_node": {"ordinary_tick": {"dram_bytes": mean_o, "seconds": sec_o, "gbs": mean_o / sec_o / 1e9, "ticks": len(ordinary)}}}
if has_cycle:
    out["k_node"]["dram_bytes_per_launch"] = (19 * mean_o + big["dram_bytes"]) / 20
    out["k_node"]["docycle_tick"] = {"dram_bytes": big["dram_bytes"], "seconds": big["seconds"], "gbs": big["dram_bytes"] / big["seconds"] / 1e9}
    out["k_node"]["how"] = ("ncu --set full --clock-control none, k_node_msgs + k_node_tasks of %d consecutive ticks around t = 1600 ms; "
                            "per-tick average = (19 x ordinary + 1 x doCycle) / 20" % len(ticks))
else:
    out["k_node"]["dram_bytes_per_launch"] = mean_o
    out["k_node"]["how"] = ("ncu --set full --clock-control none, k_node_msgs + k_node_tasks of %d consecutive ORDINARY ticks from t = 1590 ms "
                            "(the capture ended before the doCycle tick of t = 1601); compare with the algorithmic bytes of an ordinary tick "
                            "(96 B x deliveries + 8 B x update words + 128 B x updates per tick, about 58 MB), not with the whole-run average" % len(ticks))
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
