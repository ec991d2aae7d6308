"""Summarise an .ncu-rep (run where ncu is installed, no GPU needed): python tools/ncu_summary.py rep [topN]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.avg.per_cycle_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("====", d.get("Kernel Name", "")[:90])
    for k in KEYS:
        if k in d:
            print(f"  {k:70s} {d[k]}")
    st = sorted(((float(v), k) for k, v in d.items() if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and v not in ("", None)), reverse=True)
    tot = sum(v for v, _ in st) or 1
    print("  stalls:", ", ".join(f"{k.split('stalled_')[1]}={100 * v / tot:.0f}%" for v, k in st[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for blk in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    name = rows[0][1][:80]
    hdr = rows[1]
    ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [(int(r[isamp] or 0), int(r[iex] or 0), i, r[ia].strip()) for i, r in enumerate(rows[2:]) if len(r) > isamp]
    tot = sum(d[0] for d in data) or 1
    ops = Counter()
    for s, e, i, t in data:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        ops[op.split(".")[0]] += e
    print("==== source:", name, "samples", tot, "sass lines", len(data), "warp-instr", sum(d[1] for d in data))
    print("  instr mix:", ", ".join(f"{k}={v}" for k, v in ops.most_common(14)))
    for s, e, i, t in sorted(data, reverse=True)[:topn]:
        print(f"  {100 * s / tot:5.1f}% ex={e:8d} line={i:5d} {t[:100]}")
