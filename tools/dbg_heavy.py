"""Debug helper: heavy-family values against the oracle on the small 5-taxon tree."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import beluga_b200 as B
from oracle.oracle import BRANCH as OBRANCH, OracleDLWGD
nw = "((A:0.4,B:0.5):0.3,(C:0.6,(D:0.2,E:0.3):0.2):0.25);"
names = list("ABCDE")
rows = [[60, 70, 5, 40, 30], [0, 0, 300, 0, 1], [120, 130, 110, 125, 115], [33, 33, 33, 1, 1], [20, 20, 20, 20, 20],
        [1, 0, 0, 0, 900], [1, 0, 0, 0, 200], [1, 0, 0, 0, 400], [0, 0, 150, 0, 1], [100, 100, 0, 0, 1], [0, 1, 0, 0, 130]]
counts = np.array(rows, dtype=np.int64)
model, data = B.DLWGD(nw, (names, counts), 0.9, 1.0, 0.8, B.BRANCH, device=0)
orc = OracleDLWGD(nw, names, counts, 0.9, 1.0, 0.8, OBRANCH, threads=4)
got = B.logpdf_(model, data)
want, pf = orc.logpdf(per_family=True)
g = data.family_logpdf()
print(data.kernel_runs())
for r, a, b in zip(rows, g, pf):
    print(r, a, b, abs(a - b) / abs(b))
ow, ox = orc.profile()
for f in range(len(rows)):
    if not np.isfinite(g[f]) or abs(g[f] - pf[f]) > 1e-9 * abs(pf[f]):
        for node in range(1, 10):
            gc = data.column(f, node); oc = orc.column(f, node)
            n = min(len(gc), len(oc))
            fin = np.isfinite(oc[:n])
            bad = np.where(~np.isclose(gc[:n][fin], oc[:n][fin], rtol=1e-9, atol=1e-9))[0]
            print("  fam", f, "node", node, "x", ox[f][node - 1], "mismatches", len(bad), bad[:5], gc[:n][fin][bad[:3]], oc[:n][fin][bad[:3]],
                  "nonfinite got", int((~np.isfinite(gc[:n][fin])).sum()))
        break
