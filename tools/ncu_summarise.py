"""Summarise an `ncu --set full` report here (no GPU needed):  python tools/ncu_summarise.py gpurun_out/x.ncu-rep [tag]

Writes profiles/<tag>_ncu_full.txt (one line per profiled launch: duration, DRAM bytes, DRAM throughput %, tensor-pipe
%, issue-slot %, registers, achieved occupancy) and merges per-kernel DRAM traffic into profiles/ncu_traffic.json, which
bench.py's roofline.traffic reads."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
tag = sys.argv[2] if len(sys.argv) > 2 else os.path.splitext(os.path.basename(rep))[0]
if rep.endswith(".csv"):       # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): the report itself was too big to bring back
    raw = open(rep).read()
    raw = raw[raw.index('"ID"'):]
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=None):
    i = col.get(name)
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return r[i]


def unit(name):
    return units[col[name]] if name in col else ""


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1)


lines, kernels = [], []
for r in data:
    name = r[col["Kernel Name"]]
    dur = to_us(get(r, "gpu__time_duration.sum", 0.0), unit("gpu__time_duration.sum"))
    rd = to_bytes(get(r, "dram__bytes_read.sum", 0.0), unit("dram__bytes_read.sum"))
    wr = to_bytes(get(r, "dram__bytes_write.sum", 0.0), unit("dram__bytes_write.sum"))
    ent = {"kernel": name, "grid": r[col["Grid Size"]] if "Grid Size" in col else "", "duration_us": dur, "dram_bytes": rd + wr,
           "dram_read": rd, "dram_write": wr,
           "dram_pct": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "tensor_pct": get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                             get(r, "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active")),
           "issue_pct": get(r, "sm__inst_issued.avg.pct_of_peak_sustained_active",
                            get(r, "smsp__issue_active.avg.pct")),
           "regs": get(r, "launch__registers_per_thread"), "occupancy_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
           "l2_hit_pct": get(r, "lts__t_sector_hit_rate.pct")}
    kernels.append(ent)
    lines.append("%-46s grid %-14s %8.1f us  DRAM %8.2f MB (R %7.2f W %7.2f) = %6.0f GB/s  dram%% %s  tensor%% %s  issue%% %s  regs %s  occ%% %s  L2hit%% %s" % (
        name[:46], ent["grid"], dur, (rd + wr) / 1e6, rd / 1e6, wr / 1e6, (rd + wr) / max(dur, 1e-9) / 1e3, ent["dram_pct"],
        ent["tensor_pct"], ent["issue_pct"], ent["regs"], ent["occupancy_pct"], ent["l2_hit_pct"]))
out = os.path.join(ROOT, "profiles", tag + "_ncu_full.txt")
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none, %s (%d launches; every second launch of a kernel is the cold-L2 one)\n" % (os.path.basename(rep), len(lines)))
    f.write("\n".join(lines) + "\n")
print("\n".join(lines))
tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
db = json.load(open(tpath)) if os.path.exists(tpath) else {"kernels": []}
db["source"] = "profiles/%s_ncu_full.txt" % tag
keep = [k for k in db["kernels"] if k.get("tag") != tag]
order = {}
for k in kernels:
    order.setdefault(k["kernel"], []).append(k)
for name, lst in order.items():
    for i, k in enumerate(lst):
        keep.append({"tag": tag, "kernel": name, "shape": "launch %d of %d in tools/ncu_kernels.py" % (i + 1, len(lst)), "grid": k["grid"],
                     "dram_bytes": k["dram_bytes"], "duration_us": k["duration_us"], "tensor_pct": k["tensor_pct"], "dram_pct": k["dram_pct"]})
db["kernels"] = keep
json.dump(db, open(tpath, "w"), indent=1)
print("wrote", out, tpath)
