"""Copy the last measurement pass (tools/final_measure.sh, tools/final_ncu.sh) from gpurun_out/ into profiles/ and
rebuild the summaries (run in the build container, ncu needed only to read the reports)."""
import collections
import csv
import glob
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for src, dst in [("final_bench.json", "r1_bench.json"), ("final_bench_reference.json", "r1_bench_reference_cpu.json"),
                 ("final_latency.json", "r1_latency.json"), ("check_insitu.txt", "r1_steps_b16_insitu.txt"),
                 ("check_steps.txt", "r1_steps_b16.txt"), ("final_launches.csv", "r1_ncu_launch_list.csv"),
                 ("bench_2gpu.json", "r1_bench_2gpu.json")]:
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

rows = [r for r in csv.reader(open(os.path.join(P, "r1_ncu_launch_list.csv"))) if len(r) > 5]
hdr, data = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
    elif hdr and r[0].isdigit():
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("hp::", ""))
    us = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(d["Metric Unit"], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
out = ["ncu launch list of `python bench.py --steps 2 --warmup 1` (first 400 launches, --clock-control none; cold-cache and",
       "serialised: only the SHARE of each kernel is comparable with the CUDA-event numbers in r1_bench.json /",
       "r1_steps_b16_insitu.txt)", "", f"{'kernel':28s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k:28s} {n:8d} {t:10.1f} {t / n:8.2f} {100 * t / tot:6.1f}%")
out.append(f"{'total':28s} {sum(a[0] for a in agg.values()):8d} {tot:10.1f}")
open(os.path.join(P, "r1_ncu_launch_list_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

reps = sorted(glob.glob(os.path.join(G, "final_ncu_*.ncu-rep")))
for rep in reps:
    k = os.path.basename(rep)[len("final_ncu_"):-len(".ncu-rep")]
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "10"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"r1_ncu_full_{k}.txt"), "w").write(txt)
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py")] + reps)
