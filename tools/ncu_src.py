"""Summarise an `ncu --page source --csv` dump: top stall sites and the instruction mix.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_src.py src.csv [top]"""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; body = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print({k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = body[i]
    st = {s[6:]: int(r[ix[s]]) for s in stalls if int(r[ix[s]])}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5d %5.2f%% %-70s %s  conf=%s" % (i, 100 * int(r[ix["# Samples"]]) / tot, r[ix["Source"]].strip()[:70], main, r[ix["L1 Wavefronts Shared Excessive"]]))
mix = {}
for r in body:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    mix[op] = mix.get(op, 0) + int(r[ix["Instructions Executed"]])
t = sum(mix.values())
print({k: round(100 * v / t, 1) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:16]})
