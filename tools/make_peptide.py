#!/usr/bin/env python
"""Build the BASELINE config-5 test molecule: a ~100-atom drug-like system with N, O, S and Cl.

Ac-Ala-Cys-Ser-Met-(4-Cl-Phe)-Gly-Ala-NHMe, an end-capped heptapeptide in an extended conformation, built from ideal internal
coordinates (no structure file in the reference fits config 5) and then relaxed with the CPU oracle's GFN2 energy/gradient
(scipy L-BFGS).  Output: entry "peptide_cl" of qcxms_b200/data/molecules.json (bohr).  Run in the build container.
"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np

AATOAU = 1.0 / 0.52917726
Z = {'H': 1, 'C': 6, 'N': 7, 'O': 8, 'S': 16, 'Cl': 17}


class Mol:
    def __init__(self):
        self.sym, self.xyz = [], []

    def first(self, s, p):
        self.sym.append(s); self.xyz.append(np.array(p, float)); return len(self.sym) - 1

    def add(self, s, a, b, c, r, ang, dih):
        """atom bonded to c: |c-new| = r, angle(b, c, new) = ang, dihedral(a, b, c, new) = dih (degrees)"""
        A, B, C = self.xyz[a], self.xyz[b], self.xyz[c]
        ang, dih = np.radians(ang), np.radians(dih)
        bc = C - B; bc /= np.linalg.norm(bc)
        n = np.cross(B - A, bc); n /= np.linalg.norm(n)
        m = np.cross(n, bc)
        d = np.array([-r * np.cos(ang), r * np.sin(ang) * np.cos(dih), r * np.sin(ang) * np.sin(dih)])
        self.sym.append(s); self.xyz.append(C + d[0] * bc + d[1] * m + d[2] * n)
        return len(self.sym) - 1


def methyl(m, a, b, c, sym='C'):
    """-CH3 on atom c (neighbours b, a)"""
    k = m.add(sym, a, b, c, 1.52, 110.0, 180.0) if sym else c
    for d in (60.0, 180.0, 300.0):
        m.add('H', b, c, k, 1.09, 109.5, d)
    return k


def build():
    m = Mol()
    PHI, PSI = -139.0, 135.0
    # acetyl cap: CH3-C(=O)-
    ch3 = m.first('C', [0.0, 0.0, 0.0])
    c = m.first('C', [1.52, 0.0, 0.0])
    o = m.first('O', [2.15, 1.06, 0.0])
    for d in (0.0, 120.0, 240.0):
        m.add('H', o, c, ch3, 1.09, 109.5, d)
    prev_ca, prev_c, prev_o = ch3, c, o
    n = m.add('N', o, prev_ca, prev_c, 1.335, 116.6, 180.0)
    for res in ['ALA', 'CYS', 'SER', 'MET', 'CLF', 'GLY', 'ALA']:
        m.add('H', prev_o, prev_c, n, 1.01, 119.0, 180.0)
        ca = m.add('C', prev_ca, prev_c, n, 1.458, 121.9, 180.0)
        c = m.add('C', prev_c, n, ca, 1.525, 111.0, PHI)
        o = m.add('O', n, ca, c, 1.231, 120.5, PSI + 180.0)
        m.add('H', c, n, ca, 1.09, 109.0, 119.0)
        if res == 'GLY':
            m.add('H', c, n, ca, 1.09, 109.0, -119.0)
        else:
            cb = m.add('C', c, n, ca, 1.53, 110.5, -122.5)
            if res == 'ALA':
                for d in (60.0, 180.0, 300.0):
                    m.add('H', n, ca, cb, 1.09, 109.5, d)
            else:
                m.add('H', n, ca, cb, 1.09, 109.5, 60.0)
                m.add('H', n, ca, cb, 1.09, 109.5, -60.0)
                if res == 'CYS':
                    sg = m.add('S', n, ca, cb, 1.81, 114.0, 180.0)
                    m.add('H', ca, cb, sg, 1.34, 96.0, 180.0)
                elif res == 'SER':
                    og = m.add('O', n, ca, cb, 1.42, 111.0, 180.0)
                    m.add('H', ca, cb, og, 0.96, 108.0, 180.0)
                elif res == 'MET':
                    cg = m.add('C', n, ca, cb, 1.53, 113.0, 180.0)
                    m.add('H', ca, cb, cg, 1.09, 109.5, 60.0)
                    m.add('H', ca, cb, cg, 1.09, 109.5, -60.0)
                    sd = m.add('S', ca, cb, cg, 1.81, 112.0, 180.0)
                    ce = m.add('C', cb, cg, sd, 1.79, 100.0, 180.0)
                    for d in (60.0, 180.0, 300.0):
                        m.add('H', cg, sd, ce, 1.09, 109.5, d)
                elif res == 'CLF':
                    cg = m.add('C', n, ca, cb, 1.51, 113.0, 180.0)
                    cd1 = m.add('C', ca, cb, cg, 1.39, 120.0, 90.0)
                    cd2 = m.add('C', ca, cb, cg, 1.39, 120.0, -90.0)
                    ce1 = m.add('C', cb, cg, cd1, 1.39, 120.0, 180.0)
                    ce2 = m.add('C', cb, cg, cd2, 1.39, 120.0, 180.0)
                    cz = m.add('C', cg, cd1, ce1, 1.39, 120.0, 0.0)
                    m.add('H', cb, cg, cd1, 1.08, 120.0, 0.0)
                    m.add('H', cb, cg, cd2, 1.08, 120.0, 0.0)
                    m.add('H', cg, cd1, ce1, 1.08, 120.0, 180.0)
                    m.add('H', cg, cd2, ce2, 1.08, 120.0, 180.0)
                    m.add('Cl', cd1, ce1, cz, 1.74, 120.0, 180.0)
        prev_ca, prev_c, prev_o = ca, c, o
        n = m.add('N', n, ca, c, 1.335, 116.6, PSI)
    # N-methyl amide cap
    m.add('H', prev_o, prev_c, n, 1.01, 119.0, 180.0)
    cm = m.add('C', prev_ca, prev_c, n, 1.458, 121.9, 180.0)
    for d in (60.0, 180.0, 300.0):
        m.add('H', prev_c, n, cm, 1.09, 109.5, d)
    return m


def main():
    from scipy.optimize import minimize
    from oracle import pyoracle as po
    m = build()
    num = np.array([Z[s] for s in m.sym], dtype=np.int32)
    xyz = np.array(m.xyz) * AATOAU
    d = np.linalg.norm(xyz[:, None] - xyz[None], axis=2) + 10 * np.eye(len(num))
    print("atoms", len(num), "min distance / bohr", d.min(), "nao", po.dims(num))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150

    def fun(x):
        r = po.egrad(num, x.reshape(-1, 3), charge=0, multiplicity=1, etemp=300.0)
        assert r["stat"] == 0
        return r["energy"], r["gradient"].ravel()

    res = minimize(fun, xyz.ravel(), jac=True, method="L-BFGS-B", options={"maxiter": steps, "gtol": 2e-4})
    xyz = res.x.reshape(-1, 3)
    e, g = fun(res.x)
    print("E = %.8f Eh, max |g| = %.2e Eh/bohr after %d iterations" % (e, np.abs(g).max(), res.nit))
    path = os.path.join(os.path.dirname(__file__), "..", "qcxms_b200", "data", "molecules.json")
    db = json.load(open(path))
    db["peptide_cl"] = {"num": [int(v) for v in num], "xyz": [[float(v) for v in r] for r in xyz], "charge": 1,
                        "note": "Ac-Ala-Cys-Ser-Met-(4-Cl-Phe)-Gly-Ala-NHMe, GFN2 (oracle) relaxed neutral geometry; BASELINE config 5"}
    json.dump(db, open(path, "w"))


if __name__ == "__main__":
    main()
