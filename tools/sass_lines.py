"""Static SASS instruction count per CUDA source line of one kernel (no GPU needed):
    python tools/sass_lines.py score_kernels KERNEL_SUBSTRING [first_line last_line]
extracts the cubin of pyrodigal_b200/csrc/build/<name>.o, disassembles it with line info (nvdisasm -g) and prints how many
instructions of the kernel are attributed to each source line (optionally only lines in [first, last])."""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, kern = sys.argv[1], sys.argv[2]
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 1 << 30)
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "pyrodigal_b200", "csrc", "build", name + ".o")], cwd=td,
                   check=True, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(td, cubin)], capture_output=True, text=True).stdout
cur_fn, cur = None, None
cnt = collections.Counter()
total = 0
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur_fn and kern in cur_fn and re.search(r"/\*[0-9a-f]{4,}\*/", line) and cur:
        total += 1
        cnt[cur] += 1
print(f"{kern}: {total} instructions")
sel = [(k, v) for k, v in cnt.items() if k[0].endswith(".cu") and lo <= k[1] <= hi or (lo == 0 and hi == 1 << 30)]
for (f, ln), v in sorted(sel, key=lambda x: -x[1])[:40] if lo == 0 else sorted(sel):
    print(f"{v:5d}  {f}:{ln}")
if lo:
    print("sum in range:", sum(v for (f, ln), v in sel))
