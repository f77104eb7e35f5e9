"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + instruction mix / stall
samples per source line.  Usage: python tools/ncu_summary.py REPORT.ncu-rep [frames_per_launch]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
units_per_launch = float(sys.argv[2]) if len(sys.argv) > 2 else None

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name", "?"))
    for k in KEYS:
        if k in d:
            print(f"  {k} [{units[hdr.index(k)]}] = {d[k]}")
    st = sorted(((float(d[h]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in STALL), reverse=True)
    print("  stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
    if units_per_launch:
        print(f"  warp-inst per unit = {float(d['smsp__inst_executed.sum']) / units_per_launch:.1f}")
        tr = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"]))
        u = units[hdr.index("dram__bytes_read.sum")]
        print(f"  dram traffic per unit = {tr / units_per_launch * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[u]:.1f} B")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
isrc = [i for i, x in enumerate(hdr) if x == "Source"]
icu, isass = isrc[0], isrc[1]
ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
byop, sampop, byline, sampline = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
tot = totsamp = 0
for r in rows[h + 1:]:
    if len(r) <= ii or not r[ii].strip().isdigit():
        continue
    n, s = int(r[ii]), (int(r[isamp]) if r[isamp].strip().isdigit() else 0)
    if r[isass].strip() not in ('', '-'):
        tok = r[isass].split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        op = op.split(".")[0]
        byop[op] += n; sampop[op] += s; tot += n; totsamp += s
    else:
        byline[r[icu].strip()[:100]] += n; sampline[r[icu].strip()[:100]] += s
scale = units_per_launch or 1.0
print(f"\ninstruction mix (warp-inst{' per unit' if units_per_launch else ''}; total {tot / scale:.1f}; samples {totsamp})")
for op, n in byop.most_common(24):
    print(f"  {op:8s} {n / scale:10.1f} {100 * n / tot:5.1f}%  samples {100 * sampop[op] / max(totsamp, 1):5.1f}%")
print("\nhottest source lines (warp-inst, % samples)")
for l, n in sorted(byline.items(), key=lambda kv: -sampline[kv[0]])[:28]:
    print(f"  {n / scale:9.1f} {100 * sampline[l] / max(totsamp, 1):5.1f}%  {l}")
