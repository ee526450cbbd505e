"""Attribute an ncu source-page CSV (SASS view) to CUDA source lines using nvdisasm -g line info.
usage: sass_by_line.py src.csv cubin mangled_kernel_name [top_n]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kern = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":"))
line_of = {}; cur = None
for l in dis[start + 1:]:
    if l.startswith("//---------------------") or l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
base = int(body[0][col["Address"]], 16)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
tot_e = tot_s = 0.0
for r in body:
    off = int(r[col["Address"]], 16) - base
    key = line_of.get(off)
    e = float(r[col["Instructions Executed"]]); s = float(r[col["# Samples"]])
    a = agg[key[:2] if key else None]; a[0] += e; a[1] += s; a[2] += 1
    tot_e += e; tot_s += s
srcs = {}
def text(f, n):
    if f not in srcs:
        try: srcs[f] = open(f"/root/repo/raytrace_clj_b200/csrc/{f}").read().splitlines()
        except Exception: srcs[f] = []
    return srcs[f][n - 1].strip()[:90] if 0 < n <= len(srcs[f]) else ""
print(f"executed warp instructions {tot_e:.0f}, samples {tot_s:.0f}")
for key, (e, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    f, ln = key if key else ("?", 0)
    print(f"{100 * e / tot_e:5.1f}% exec {100 * s / tot_s:5.1f}% samples {n:4d} instr  {f}:{ln}  {text(f, ln)}")
