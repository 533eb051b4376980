"""Aggregate the stall samples and executed instructions of one kernel in an .ncu-rep by CUDA source line.
ncu's CSV source page is per SASS instruction; nvdisasm --print-line-info of the same build gives the line of every
instruction, in the same order.
usage: ncu_by_line.py <report.ncu-rep> <kernel symbol substring>[:<substring of its mangled name>] [top N]   (run where the library was built)"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, sym = sys.argv[1], sys.argv[2]
sec = sym                                   # "name:mangled" when the substring that picks the kernel in the report does not pick its ELF section
if ":" in sym: sym, sec = sym.split(":", 1)
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "lzbench_b200/_lib/liblz4b200.so")], cwd=tmp, check=True,
               stdout=subprocess.DEVNULL)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
lines = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(lines) if ".section" in l and ".text." in l and sec in l)
end = next((i for i in range(start + 5, len(lines)) if ".section" in lines[i]), len(lines))
cur, seq = None, []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+.+?;", l):
        seq.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# a report may hold several kernels: every one starts with a "Kernel Name" row followed by the column header
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
pick = next(i for i in starts if sym in rows[i][1])
stop = next((i for i in starts if i > pick), len(rows))
hdr, data = rows[pick + 1], [r for r in rows[pick + 2:stop] if r]
ia, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(seq) == len(data), (len(seq), len(data), "the report is from another build")
by = collections.defaultdict(lambda: [0, 0])
for k, r in zip(seq, data):
    by[k][0] += int(r[ismp]); by[k][1] += int(r[ia])
ts, ti = sum(v[0] for v in by.values()), sum(v[1] for v in by.values())
src = {}
print(f"{ts} samples, {ti} warp instructions")
for k, v in sorted(by.items(), key=lambda x: -x[1][0])[:top]:
    f, n = k if k else ("?", 0)
    if f not in src:
        p = os.path.join(ROOT, "lzbench_b200/csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][n - 1].strip()[:100] if 0 < n <= len(src[f]) else ""
    print(f"{100 * v[0] / ts:5.1f} % samples {100 * v[1] / ti:5.1f} % instr  {f}:{n}  {text}")
