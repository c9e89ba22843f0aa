"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep captured with --import-source on
(not a pytest file).  usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [min_pct]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, fname, h = [], None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": h = r; continue
    if h is None or len(r) < len(h) or not r[0]: continue
    ie, sm = h.index("Instructions Executed"), h.index("# Samples")
    if r[ie].isdigit() and r[sm].isdigit(): lines.append((fname, int(r[0]), r[1].strip(), int(r[ie]), int(r[sm])))
ti, ts = sum(l[3] for l in lines), sum(l[4] for l in lines)
print(f"total warp instructions {ti}, samples {ts}")
for f, n, s, i, smp in lines:
    if i >= ti * min_pct / 100 or smp >= ts * min_pct / 100:
        print(f"{f}:{n:<4d} inst {100*i/ti:5.1f}%  samples {100*smp/max(ts,1):5.1f}%  {s[:120]}")
