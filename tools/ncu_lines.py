"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by CUDA source line:
samples, instructions executed and the leading stall reasons per line."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = None
agg = collections.OrderedDict()
cur_line = None
tot_s = tot_i = 0
stall_tot = collections.Counter()
for r in rows:
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        ix = {name: i for i, name in enumerate(hdr)}
        # the second "Source" column is the SASS text
        sass_col = [i for i, n in enumerate(hdr) if n == "Source"][1]
        stalls = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0].strip():
        cur_line = (int(r[0]), r[1].strip()[:110])
    if not r[sass_col].strip():
        continue
    try:
        smp = int(r[ix["# Samples"]] or 0); ins = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    a = agg.setdefault(cur_line, {"s": 0, "i": 0, "st": collections.Counter()})
    a["s"] += smp; a["i"] += ins
    tot_s += smp; tot_i += ins
    for i, n in stalls:
        v = int(r[i] or 0)
        a["st"][n] += v; stall_tot[n] += v
print("total samples %d  warp instructions %d" % (tot_s, tot_i))
print("stalls:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / max(tot_s, 1)) for n, v in stall_tot.most_common(8)))
for (ln, txt), a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:top]:
    st = ", ".join("%s %d" % (n[6:], v) for n, v in a["st"].most_common(3))
    print("%5d  smp %5.1f%%  ins %5.1f%%  [%s]  %s" % (ln, 100.0 * a["s"] / tot_s, 100.0 * a["i"] / max(tot_i, 1), st, txt))
