"""Per-source-line summary of one kernel of an ncu report (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top]
prints instructions executed / stall samples per CUDA source line (cuda,sass correlated source page)."""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; data = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ie = r.index("Instructions Executed"); sm = r.index("# Samples"); te = r.index("Thread Instructions Executed"); continue
    if hdr and r[0].isdigit():
        try: data.append((int(r[ie]), int(r[sm] or 0), int(r[te]), cur_file, int(r[0]), r[1].strip()))
        except ValueError: pass
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print(f"total warp instructions {tot:,}  samples {tots:,}")
for v, s, t, f, ln, src in sorted(data, reverse=True)[:top]:
    print(f"{100*v/tot:5.1f}% inst {100*s/max(tots,1):5.1f}% smp  lanes {t/max(v,1):4.1f}  {f}:{ln}: {src[:100]}")
