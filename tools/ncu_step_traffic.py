"""Per-kernel DRAM traffic of ONE WHOLE STEP from the ncu metrics pass of tools/r2_profile.sh (every launch of the step,
not hand-picked ones): python tools/ncu_step_traffic.py gpurun_out/r2_step_traffic.csv -> profiles/ncu_traffic.json + a
launch list under profiles/."""
import csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
launches = {}
for r in rows:
    e = launches.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        e["duration_us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    else:
        e[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
out = {}
tot_us = sum(e["duration_us"] for e in launches.values())
for i in sorted(launches):
    e = launches[i]
    short = re.sub(r"<.*", "", re.sub(r"\(.*", "", e["kernel"]).replace("void ", "").strip()).replace("hp::", "")
    k = out.setdefault(short, {"launches": 0, "duration_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    k["launches"] += 1
    k["duration_us"] += e["duration_us"]
    k["dram_read_bytes"] += e.get("dram__bytes_read.sum", 0.0)
    k["dram_write_bytes"] += e.get("dram__bytes_write.sum", 0.0)
for k, v in out.items():
    v["dram_bytes_per_launch"] = round((v["dram_read_bytes"] + v["dram_write_bytes"]) / v["launches"])
    v["share_of_step_time"] = round(v["duration_us"] / tot_us, 4)
    v["duration_us"] = round(v["duration_us"], 2)
    v["source"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over ALL "
                   f"launches of one step (batch 16, fast mode, 4th step of tools/one_step.py): {os.path.basename(path)}")
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
with open(os.path.join(ROOT, "profiles", "r2_ncu_step_launch_list.csv"), "w") as f:
    f.write("id,kernel,grid,block,duration_us,dram_read_bytes,dram_write_bytes\n")
    for i in sorted(launches):
        e = launches[i]
        name = re.sub(r"\(.*", "", e["kernel"]).replace("void ", "").strip()
        f.write(f'{i},"{name}","{e["grid"]}","{e["block"]}",{e["duration_us"]:.2f},{e.get("dram__bytes_read.sum", 0):.0f},{e.get("dram__bytes_write.sum", 0):.0f}\n')
print(f"{len(launches)} launches, {tot_us:.1f} us (cold-cache, serialised)")
for k, v in sorted(out.items(), key=lambda kv: -kv[1]["duration_us"]):
    print(f'{k:24s} {v["launches"]:3d} launches {v["duration_us"]:8.1f} us ({100 * v["share_of_step_time"]:4.1f} %)  DRAM {v["dram_bytes_per_launch"] / 1e6:7.2f} MB/launch')
