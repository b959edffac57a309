"""Turn an ncu launch log (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
of `bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-step` into the per-launch table of ONE layer
step (the first timed step) and its DRAM-traffic sum.
usage: python tools/launch_list.py gpurun_out/launches.csv profiles/r1_launches_v3_general.csv [nth_step]"""
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 4           # 3 warm-up steps precede the first timed one
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    lid = int(r[col["ID"]])
    d = launches.setdefault(lid, {"name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
    val = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    name = r[col["Metric Name"]]
    if name == "gpu__time_duration.sum":
        d["us"] = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    else:
        mb = val * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
        d["rd" if "read" in name else "wr"] = mb
ids = sorted(launches)
short = lambda n: re.sub(r"\(.*", "", re.sub(r"^void ", "", n)).replace("x2::", "")
fwd = [i for i in ids if "k_attn_fwd" in launches[i]["name"]]
a = fwd[nth - 1]
# a step starts at the k_rbf_filter before this forward attention and ends before the next step's k_rbf_filter pair
beg = max(i for i in ids if i < a and "k_rbf_filter" in launches[i]["name"])
nxt = [i for i in fwd if i > a]
end = max(i for i in ids if i < nxt[0] and "k_rbf_filter" in launches[i]["name"]) if nxt else ids[-1] + 1
step = [i for i in ids if beg <= i < end]
tot = sum(launches[i]["us"] for i in step)
with open(dst, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ; "
            "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-step\n")
    f.write("# one SBFTransformerConv fwd+bwd step (headline workload, X2_MODE_TF32X3); launch order; ncu times are "
            "cold-cache/serialised: compare shares\n")
    f.write("kernel,grid,block,time_us,share_pct,dram_read_MB,dram_write_MB\n")
    for i in step:
        d = launches[i]
        f.write(f'{short(d["name"])},"{d["grid"]}","{d["block"]}",{d["us"]:.1f},{100 * d["us"] / tot:.1f},'
                f'{d.get("rd", 0):.1f},{d.get("wr", 0):.1f}\n')
    rd, wr = sum(launches[i].get("rd", 0) for i in step), sum(launches[i].get("wr", 0) for i in step)
    f.write(f"# total,{len(step)} launches,,{tot:.1f},100.0,{rd:.1f},{wr:.1f}\n")
print(json.dumps({"launches": len(step), "time_us": round(tot, 1), "dram_read_bytes": rd * 1e6, "dram_write_bytes": wr * 1e6,
                  "dram_bytes_per_step": int((rd + wr) * 1e6)}))
