"""Extract per-launch DRAM traffic and durations from `ncu --set full` reports into profiles/ncu_traffic.json.

    python tools/ncu_traffic.py gpurun_out/ncu_*.ncu-rep
"""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").strip()
        short = re.sub(r"<.*", "", name).replace("hp::", "")
        unit = lambda k: (d.get(k) or "0").replace(",", "")
        # raw page reports bytes in the unit of the second header row
        units = dict(zip(hdr, rows[1]))
        def to_bytes(k):
            v = float(unit(k)); u = units.get(k, "byte")
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        def to_us(k):
            v = float(unit(k)); u = units.get(k, "ns")
            return v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        e = out.setdefault(short, {"launches": []})
        e["launches"].append({"kernel": name, "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size"),
                              "duration_us": round(to_us("gpu__time_duration.sum"), 2),
                              "dram_read_bytes": to_bytes("dram__bytes_read.sum"),
                              "dram_write_bytes": to_bytes("dram__bytes_write.sum"),
                              "issue_active_pct": float(unit("sm__issue_active.avg.pct_of_peak_sustained_elapsed")),
                              "warps_active_pct": float(unit("sm__warps_active.avg.pct_of_peak_sustained_active")),
                              "registers": d.get("launch__registers_per_thread"), "report": os.path.basename(rep)})
for k, e in out.items():
    L = e["launches"]
    e["dram_bytes_per_launch"] = round(sum(x["dram_read_bytes"] + x["dram_write_bytes"] for x in L) / len(L))
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in out.items()}))
