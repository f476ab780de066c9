#!/usr/bin/env python
"""Build qcxms_b200/data/molecules.json: the benchmark/test input geometries.

* the four share/examples molecules of the reference (coord files, bohr) -- benchmark inputs, SURVEY.md 8(d);
* caffeine (BASELINE config 2): not in the reference; a standard 3-D structure relaxed here with the CPU oracle
  (scipy L-BFGS on the oracle's GFN2 energy/gradient) to a GFN2 minimum.
Run in the build container only (needs /root/reference).
"""
import json, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from scipy.optimize import minimize
from oracle import pyoracle as po

SYM = {'h': 1, 'he': 2, 'c': 6, 'n': 7, 'o': 8, 'f': 9, 's': 16, 'cl': 17, 'ar': 18}
AATOAU = 1.0 / 0.52917726

def read_coord(path):
    num, xyz = [], []
    for line in open(path):
        p = line.split()
        if len(p) == 4 and p[3].lower() in SYM:
            xyz.append([float(v) for v in p[:3]]); num.append(SYM[p[3].lower()])
    return num, xyz

CAFFEINE_AA = """O 0.4700 2.5688 0.0006
O -3.1271 -0.4436 -0.0003
N -0.9686 -1.3125 0.0000
N 2.2182 0.1412 -0.0003
N -1.3477 1.0797 -0.0001
N 1.4119 -1.9372 0.0002
C 0.8579 0.2592 -0.0008
C 0.3897 -1.0264 -0.0004
C 0.0307 1.4220 -0.0006
C -1.9061 -0.2495 -0.0004
C 2.5032 -1.1998 0.0003
C -1.4276 -2.6960 0.0008
C 3.1926 1.2061 0.0003
C -2.2969 2.1881 0.0007
H 3.5163 -1.5787 0.0008
H -1.0451 -3.1973 -0.8937
H -2.5186 -2.7596 0.0011
H -1.0447 -3.1963 0.8957
H 4.1992 0.7801 0.0002
H 3.0468 1.8092 -0.8992
H 3.0466 1.8083 0.9004
H -1.8087 3.1651 -0.0003
H -2.9322 2.1027 0.8881
H -2.9346 2.1021 -0.8849"""

def relax(num, xyz, charge=0):
    po.set_accuracy(0.01)
    n = len(num)
    def f(x):
        r = po.egrad(num, x.reshape(n, 3), charge=charge, multiplicity=1, etemp=300.0)
        return r["energy"], r["gradient"].ravel()
    res = minimize(f, np.array(xyz).ravel(), jac=True, method="L-BFGS-B", options=dict(maxiter=2000, gtol=2e-5, ftol=1e-14))
    po.set_accuracy(1.0)
    e, g = f(res.x)
    print("relaxed: E=%.8f max|g|=%.2e iterations=%d" % (e, abs(g).max(), res.nit), file=sys.stderr)
    return res.x.reshape(n, 3).tolist()

out = {}
ref = "/root/reference/share/examples/"
for name, path, chg in (("chloroethanol", "EI/2-Chloroethanol_GFN2", 0), ("monoethanolamine", "EI/Monoethanolamine_GFN1", 0),
                        ("thf_h", "CID/Tetrahydrofuran", 1), ("dichlorobenzamide_h", "CID/Dichlorobenzamide", 1)):
    num, xyz = read_coord(ref + path + "/coord")
    out[name] = dict(num=num, xyz=xyz, charge=chg, source="reference share/examples/%s/coord" % path)
lines = [l.split() for l in CAFFEINE_AA.splitlines()]
num = [SYM[l[0].lower()] for l in lines]
xyz = (np.array([[float(v) for v in l[1:]] for l in lines]) * AATOAU)
out["caffeine"] = dict(num=num, xyz=relax(num, xyz), charge=0, source="standard 3-D structure relaxed with oracle GFN2 (tools/make_molecules.py)")
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "qcxms_b200", "data", "molecules.json"), "w"), indent=1)


def alkane(nc):
    """all-trans n-alkane C_n H_{2n+2} in idealised geometry (bohr): the ~100-atom stand-in of BASELINE config 5."""
    rcc, rch, th = 1.53 * AATOAU, 1.10 * AATOAU, np.deg2rad(111.6)
    dx, dz = rcc * np.sin(th / 2), rcc * np.cos(th / 2)
    num, xyz = [], []
    cs = [np.array([i * dx, 0.0, (i % 2) * dz]) for i in range(nc)]
    for i, c in enumerate(cs):
        num.append(6); xyz.append(c)
    hy, hz = rch * np.sin(np.deg2rad(109.5) / 2), rch * np.cos(np.deg2rad(109.5) / 2)
    for i, c in enumerate(cs):
        sgn = -1.0 if i % 2 == 0 else 1.0
        for s in (-1.0, 1.0):
            num.append(1); xyz.append(c + np.array([0.0, s * hy, sgn * hz]))
    for end, c in ((-1.0, cs[0]), (1.0, cs[-1])):     # terminal hydrogens along the chain direction
        sgn = 1.0 if (0 if end < 0 else nc - 1) % 2 == 0 else -1.0
        num.append(1); xyz.append(c + rch * np.array([end * np.sin(th / 2), 0.0, sgn * np.cos(th / 2)]))
    return num, np.array(xyz)


if "--alkane-only" in sys.argv:
    path = os.path.join(os.path.dirname(__file__), "..", "qcxms_b200", "data", "molecules.json")
    out = json.load(open(path))
for _nc in (14, 17):      # medium bases between the strip-GEMM limit (72 AOs) and the shared-memory limit (~110)
    num, xyz = alkane(_nc)
    out["alkane_c%d" % _nc] = dict(num=num, xyz=xyz.tolist(), charge=0, source="idealised all-trans C%dH%d (%d atoms, %d AOs), tools/make_molecules.py" % (_nc, 2 * _nc + 2, 3 * _nc + 2, 6 * _nc + 2))
num, xyz = alkane(32)
out["alkane_c32"] = dict(num=num, xyz=xyz.tolist(), charge=0, source="idealised all-trans C32H66 (98 atoms, 194 AOs), tools/make_molecules.py")
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "qcxms_b200", "data", "molecules.json"), "w"), indent=1)
