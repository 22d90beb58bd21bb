"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv): top stalled SASS instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# first kernel section only (SASS view): header at row 1
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if r and r[0] == "Address":
        break
    if len(r) == len(hdr):
        body.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in body)
print("total samples", tot, "instructions", len(body))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in body) for h in stall_cols}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
body_sorted = sorted(enumerate(body), key=lambda t: -int(t[1][ix["# Samples"]]))
for pos, r in body_sorted[:topn]:
    reasons = {h[6:]: int(r[ix[h]]) for h in stall_cols if int(r[ix[h]])}
    print(f"{pos:5d} {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}%  {r[ix['Source']].strip():60s} {reasons}")
