"""Reverse-mode (adjoint sweep) vs forward-mode (tangent passes) gradient of the same library, same inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import beluga_b200 as B

DATA = os.path.join(os.path.dirname(__file__), "..", "tests", "data")


def read_csv(name):
    rows = [ln.strip().split(",") for ln in open(os.path.join(DATA, name))]
    return rows[0], np.array(rows[1:], dtype=np.int64)


def run(mode, nw, df, lam, mu, eta, nt, wgds=(), wgts=(), cr=False, hetero=True):
    os.environ["BLG_GRAD"] = mode
    model, data = B.DLWGD(nw, df, lam, mu, eta, nt, device=0)
    rng = np.random.default_rng(3)
    if hetero:
        for n in model.nodes.values():
            if "λ" in n.theta:
                n["λ"] = lam * float(np.exp(rng.normal(0, 0.3))); n["μ"] = mu * float(np.exp(rng.normal(0, 0.3)))
    for i, t, q in wgds:
        model.addwgd_(model[i], t, q); B.extend_(data, i)
    for i, t, q in wgts:
        model.addwgt_(model[i], t, q); B.extend_(data, i)
    t0 = time.perf_counter()
    g = B.gradient_cr(model, data) if cr else B.gradient(model, data)
    return np.asarray(g), time.perf_counter() - t0


cases = [
    ("9dicots Branch", "9dicots.nw", "9dicots-f01-25.csv", 1.0, 0.92, 0.66, B.BRANCH, (), (), False),
    ("9dicots Node", "9dicots.nw", "9dicots-f01-25.csv", 1.0, 0.92, 0.66, B.NODE, (), (), False),
    ("9dicots Branch + 3 WGD", "9dicots.nw", "9dicots-f01-25.csv", 1.0, 0.92, 0.66, B.BRANCH, ((6, 0.01, 0.25), (14, 0.02, 0.5), (10, 0.05, 0.3)), (), False),
    ("9dicots Node + WGD + WGT", "9dicots.nw", "9dicots-f01-25.csv", 0.8, 0.9, 0.7, B.NODE, ((6, 0.01, 0.25),), ((12, 0.03, 0.4),), False),
    ("12monocots Branch", "12monocots.nw", "12monocots-f01-250.csv", 0.004, 0.003, 0.9, B.BRANCH, (), (), False),
    ("12monocots cr", "12monocots.nw", "12monocots-f01-250.csv", 0.004, 0.003, 0.9, B.BRANCH, (), (), True),
]
bad = 0
for name, nwf, csv, lam, mu, eta, nt, wgds, wgts, cr in cases:
    nw = open(os.path.join(DATA, nwf)).readline().strip()
    df = read_csv(csv)
    gf, tf = run("forward", nw, df, lam, mu, eta, nt, wgds, wgts, cr, hetero=not cr)
    ga, ta = run("adjoint", nw, df, lam, mu, eta, nt, wgds, wgts, cr, hetero=not cr)
    scale = np.abs(gf).max()
    err = np.abs(ga - gf).max() / scale
    print("%-28s P=%3d  max|Δ|/max|g| = %.2e   forward %.1f ms  adjoint %.1f ms" % (name, len(gf), err, tf * 1e3, ta * 1e3))
    if not (err < 1e-9):
        bad += 1
        k = int(np.abs(ga - gf).argmax())
        print("   worst component", k, gf[k], ga[k])
        print("   forward", np.array2string(gf, precision=5, max_line_width=200))
        print("   adjoint", np.array2string(ga, precision=5, max_line_width=200))
sys.exit(1 if bad else 0)
