import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import beluga_b200 as B
from conftest import *
from oracle.oracle import NODE as ONODE, OracleDLWGD
nw = read_nw("plants1.nw"); df = read_csv("plants1-100.csv")
lam, mu, eta = LAM3[0], MU3[0], ETA3[0]
model, data = B.DLWGD(nw, df, lam, mu, eta, B.NODE, device=0)
orc = OracleDLWGD(nw, df[0], df[1], lam, mu, eta, ONODE)
print(B.logpdf_(model, data), orc.logpdf())
for j, t, q in [(10, 0.02, 0.3), (7, 0.1, 0.6)]:
    w = B.addwgd_(model, model[j], t, q); orc.addwgd(j, t, q)
    B.extend_(data, j); orc.extend(j)
    g = B.logpdf_(model[j], data); o = orc.logpdf_from(j)
    print("add above", j, g, o)
    n = model[j]
    while n is not None:
        if n.c:
            for f in (0, 3):
                a = data.column(f, n.i); b = orc.column(f, n.i)
                m = min(len(a), len(b))
                fin = np.isfinite(b[:m])
                err = np.abs(a[:m][fin]-b[:m][fin]).max() if fin.any() else 0
                print("   node", n.i, n.kind, [c.i for c in n.c], "fam", f, "maxerr", err, "" if err < 1e-9 else (a[:m], b[:m]))
        n = n.p
