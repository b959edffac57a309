"""Top stall locations of an ncu report's source page (SASS view): `python tools/ncu_top.py file.ncu-rep [n]`."""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = rows[hi + 1:]
col = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
execd = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
print("total samples", tot, "warp-instructions", execd)
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    st = {k[6:]: int(r[col[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and int(r[col[k]] or 0) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {int(r[col['# Samples']]):6d} {100*int(r[col['# Samples']])/tot:5.1f}%  ex={r[col['Instructions Executed']]:>8}  {r[col['Source']].strip()[:70]:70s} {top}")
