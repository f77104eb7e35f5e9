"""Bucket the SASS of an .ncu-rep by kernel phase (delimited by BAR.SYNC) and print, per phase,
executed warp-instructions, stall samples and the dominant stall reasons.
Usage: python tools/ncu_phases.py REPORT.ncu-rep [units_per_launch]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
phase, phases = 0, collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    sass = r[col["Source"]].strip()
    p = phases.setdefault(phase, dict(inst=0, samp=0, n=0, st=collections.Counter(), ops=collections.Counter()))
    n = int(r[col["Instructions Executed"]] or 0)
    p["inst"] += n
    p["samp"] += int(r[col["# Samples"]] or 0)
    p["n"] += 1
    tok = sass.split()
    op = (tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")).split(".")[0]
    p["ops"][op] += n
    for s in stall_cols:
        p["st"][s] += int(r[col[s]] or 0)
    if "BAR.SYNC" in sass:
        phase += 1
tot_s = sum(p["samp"] for p in phases.values())
tot_i = sum(p["inst"] for p in phases.values())
print(f"total warp-inst/unit {tot_i / units:.1f}, samples {tot_s}")
for k, p in phases.items():
    st = ", ".join(f"{n.replace('stall_', '')}={100 * v / max(p['samp'], 1):.0f}%" for n, v in p["st"].most_common(5))
    ops = ", ".join(f"{o}:{v / units:.1f}" for o, v in p["ops"].most_common(7))
    print(f"phase {k}: sass {p['n']:5d}  inst/unit {p['inst'] / units:7.1f} ({100 * p['inst'] / tot_i:4.1f}%)  samples {100 * p['samp'] / tot_s:5.1f}%  | {st}\n          {ops}")
