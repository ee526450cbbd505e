"""Summarise an `ncu --page source --csv` export: instructions grouped into regions of equal execution
count, with opcode mix and stall-sample breakdown.  usage: sass_hot.py src.csv [min_share]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_exec = sum(float(r[col["Instructions Executed"]]) for r in body)
tot_samp = sum(float(r[col["# Samples"]]) for r in body)
print(f"instructions in kernel {len(body)}  executed {tot_exec:.0f}  samples {tot_samp:.0f}")
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
# contiguous regions with the same execution count
regions = []; cur = None
for i, r in enumerate(body):
    ex = float(r[col["Instructions Executed"]])
    if cur is None or ex != cur["ex"]:
        cur = {"ex": ex, "start": i, "rows": []}; regions.append(cur)
    cur["rows"].append(r)
for g in regions:
    ex = g["ex"] * len(g["rows"]); sm = sum(float(r[col["# Samples"]]) for r in g["rows"])
    if ex / tot_exec < min_share and sm / tot_samp < min_share: continue
    ops = collections.Counter(r[col["Source"]].split()[0 if not r[col["Source"]].strip().startswith("@") else 1].split(".")[0] for r in g["rows"])
    st = collections.Counter()
    for r in g["rows"]:
        for s in stall_cols: st[s[6:]] += float(r[col[s]])
    tops = " ".join(f"{k}={v / max(sm, 1):.0%}" for k, v in st.most_common(5))
    print(f"@{g['start']:5d} n={len(g['rows']):4d} exec/instr={g['ex']:.0f} exec={ex / tot_exec:6.1%} samples={sm / tot_samp:6.1%}  {dict(ops.most_common(8))}  {tops}")
