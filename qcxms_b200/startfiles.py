"""The file contract around the production run: start.xyz + qcxms.start of one trajectory directory TMPQCXMS/TMP.<n>.

* start.xyz    plain xyz file in Angstrom (written by mctc-lib's write_structure, reference src/utility.f90:346-351)
* qcxms.start  line 1 itrj ('(i4)'), line 2 eimp ('(D22.14)', Eh), line 3 tadd ('(D22.14)', a.u. of time), then per atom
               velo(1:3), velof ('(7D22.14)') -- wrstart / rdstart, reference src/utility.f90:363-422
"""
import os

import numpy as np

AUTOAA = 0.52917726
SYMBOL = {1: "H", 2: "He", 3: "Li", 4: "Be", 5: "B", 6: "C", 7: "N", 8: "O", 9: "F", 10: "Ne", 11: "Na", 12: "Mg", 13: "Al", 14: "Si",
          15: "P", 16: "S", 17: "Cl", 18: "Ar"}
NUMBER = {v.lower(): k for k, v in SYMBOL.items()}


def fortran_d(x, w=22, d=14):
    """Fortran D<w>.<d> edit descriptor: 0.ddddddddddddddD+ee, right-justified."""
    if x == 0.0:
        s = "0." + "0" * d + "D+00"
    else:
        e = int(np.floor(np.log10(abs(x)))) + 1
        m = abs(x) / 10.0 ** e
        ms = "%.*f" % (d, m)
        if ms.startswith("1"):          # rounding carried into the leading digit
            e += 1
            ms = "%.*f" % (d, abs(x) / 10.0 ** e)
        s = ("-" if x < 0 else "") + ms + "D%+03d" % e
    return s.rjust(w)


def write_start(dirname, itrj, num, xyz_bohr, velo, velof, eimp, tadd):
    os.makedirs(dirname, exist_ok=True)
    with open(os.path.join(dirname, "start.xyz"), "w") as f:
        f.write("%d\n\n" % len(num))
        for z, r in zip(num, np.asarray(xyz_bohr) * AUTOAA):
            f.write("%-4s %20.14f %20.14f %20.14f\n" % (SYMBOL[int(z)], r[0], r[1], r[2]))
    with open(os.path.join(dirname, "qcxms.start"), "w") as f:
        f.write("%4d\n" % itrj)
        f.write(fortran_d(eimp) + "\n")
        f.write(fortran_d(tadd) + "\n")
        for v, w in zip(np.asarray(velo), np.asarray(velof)):
            f.write("".join(fortran_d(t) for t in (v[0], v[1], v[2], w)) + "\n")


def read_start(dirname):
    """-> dict(itrj, num, xyz (bohr), velo, velof, eimp, tadd)"""
    with open(os.path.join(dirname, "start.xyz")) as f:
        lines = f.read().splitlines()
    nat = int(lines[0].split()[0])
    num, xyz = [], []
    for ln in lines[2:2 + nat]:
        p = ln.split()
        num.append(NUMBER[p[0].lower()])
        xyz.append([float(v) / AUTOAA for v in p[1:4]])
    with open(os.path.join(dirname, "qcxms.start")) as f:
        rows = f.read().replace("D", "E").replace("d", "e").splitlines()
    itrj, eimp, tadd = int(rows[0].split()[0]), float(rows[1]), float(rows[2])
    vals = np.array([[float(r[22 * k:22 * (k + 1)]) for k in range(4)] for r in rows[3:3 + nat]])
    return dict(itrj=itrj, num=np.array(num, dtype=np.int32), xyz=np.array(xyz), velo=vals[:, :3].copy(), velof=vals[:, 3].copy(), eimp=eimp, tadd=tadd)
