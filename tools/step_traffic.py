"""DRAM traffic of ONE layer step from an ncu launch list, tied to the library build that produced it.

    # on the GPU box, for each mode (tf32x3 = default kernels, tf32x3_fused = fused tile forward):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 \\
        --csv --log-file gpurun_out/launches_<mode>.csv python bench.py --layer-only --steps 2 --warmup 3 --mode <mode>
    # here (or there):
    python tools/step_traffic.py <mode> gpurun_out/launches_<mode>.csv [E T]

writes profiles/r2_launches_<mode>.csv (per-launch table of the first timed step) and updates
profiles/r2_step_traffic.json[<mode>] = {dram_bytes_per_step, per-phase MB, lib_sha256 (digest of csrc/ the .so
was built from), git sha}.  bench.py reports `roofline.traffic` from that file ONLY when the digest equals the
loaded library's."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode, src = sys.argv[1], sys.argv[2]
E, T = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (43048, 811834)
nth = 4                                     # 3 warm-up steps precede the first timed one
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
col = {h: i for i, h in enumerate(rows[0])}
L = {}
for r in rows[1:]:
    d = L.setdefault(int(r[col["ID"]]), {"name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
    val, unit, name = float(r[col["Metric Value"]].replace(",", "")), r[col["Metric Unit"]], r[col["Metric Name"]]
    if name == "gpu__time_duration.sum":
        d["us"] = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    else:
        d["rd" if "read" in name else "wr"] = val * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
ids = sorted(L)
short = lambda n: re.sub(r"\(.*", "", re.sub(r"^void ", "", n)).replace("x2::", "")
filt = [i for i in ids if "k_rbf_filter" in L[i]["name"]]
# a step = forward (one k_rbf_filter) + backward (none when saved.xs is kept): steps start at every k_rbf_filter launch
beg = filt[nth - 1]
end = filt[nth] if len(filt) > nth else ids[-1] + 1
step = [i for i in ids if beg <= i < end]


def phase(name):
    n = short(name)
    if "k_tile_fwd" in n or "k_item_merge" in n:
        return "attn_fwd"
    if "k_attn_fwd" in n:
        return "attn_fwd"
    if "k_attn_bwd_tgt" in n or "k_rows_segsum" in n:
        return "attn_bwd_tgt"
    if "k_attn_bwd_src" in n:
        return "attn_bwd_src"
    return None                             # GEMM launches are attributed by position below


tot = sum(L[i]["us"] for i in step)
rd, wr = sum(L[i].get("rd", 0) for i in step), sum(L[i].get("wr", 0) for i in step)
dst = os.path.join(ROOT, "profiles", f"r2_launches_{mode}.csv")
with open(dst, "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ; "
            f"python bench.py --layer-only --steps 2 --warmup 3 --mode {mode}\n")
    f.write("# one SBFTransformerConv fwd+bwd step (headline workload); launch order; ncu times are cold-cache/serialised: compare shares\n")
    f.write("kernel,grid,block,time_us,share_pct,dram_read_MB,dram_write_MB\n")
    for i in step:
        d = L[i]
        f.write(f'{short(d["name"])},"{d["grid"]}","{d["block"]}",{d["us"]:.1f},{100 * d["us"] / tot:.1f},{d.get("rd", 0):.1f},{d.get("wr", 0):.1f}\n')
    f.write(f"# total,{len(step)} launches,,{tot:.1f},100.0,{rd:.1f},{wr:.1f}\n")
# phases by launch order: node_proj | trow_proj | attn_fwd | attn_bwd_tgt | attn_bwd_src | trow_dgrad | trow_wgrad | node_bwd
names = [short(L[i]["name"]) for i in step]
mb = [L[i].get("rd", 0) + L[i].get("wr", 0) for i in step]
ph = {k: 0.0 for k in ("node_proj", "trow_proj", "attn_fwd", "attn_bwd_tgt", "attn_bwd_src", "trow_dgrad", "trow_wgrad", "node_bwd")}
i_fwd = max(k for k, n in enumerate(names) if phase(L[step[k]]["name"]) == "attn_fwd")
i_first_attn = min(k for k, n in enumerate(names) if phase(L[step[k]]["name"]) == "attn_fwd")
i_src = max(k for k, n in enumerate(names) if "k_attn_bwd_src" in n)
for k, n in enumerate(names):
    p_ = phase(L[step[k]]["name"])
    if p_:
        ph[p_] += mb[k]
    elif k < i_first_attn:
        ph["node_proj" if k < 2 else "trow_proj"] += mb[k]
    elif k > i_src:
        after = k - i_src
        ph["trow_dgrad" if after == 1 else "trow_wgrad" if after <= 5 else "node_bwd"] += mb[k]
_libdir = os.path.join(ROOT, "x2-gnn_b200", "lib")
_stamp = os.path.join(_libdir, "libx2gnn.so.digest")         # written next to the binary by build.py
sha = open(_stamp if os.path.exists(_stamp) else os.path.join(_libdir, "libx2gnn.sha256")).read().strip()
try:
    git = subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
except Exception:
    git = None
out_path = os.path.join(ROOT, "profiles", "r2_step_traffic.json")
allm = json.load(open(out_path)) if os.path.exists(out_path) else {}
allm[mode] = {"source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the {len(step)} launches of one layer "
                        f"step (profiles/r2_launches_{mode}.csv)",
              "dram_bytes_per_step": int((rd + wr) * 1e6), "dram_read_bytes": rd * 1e6, "dram_write_bytes": wr * 1e6,
              "launches": len(step), "ncu_time_us": round(tot, 1), "E": E, "T": T, "mode": mode,
              "phase_dram_MB": {k: round(v, 1) for k, v in ph.items()}, "lib_sha256": sha, "git": git}
json.dump(allm, open(out_path, "w"), indent=1)
print(json.dumps(allm[mode]))
