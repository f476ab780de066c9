"""Fragment bookkeeping after md(): average fragment structures, fragment IPs, statistical charge assignment and the
qcxms.res records.  Host-side mirror of

  avg_frag_struc      reference src/analyse.f90:452-502
  analyse             reference src/analyse.f90:14-448        (xtb2 branch: iprog = 8, IPs with GFN2-xTB at 300 K)
  eqm / eself         reference src/iniqm.f90:241-427, 430-451 (electrons_amount :45-63, get_core_e :17-41)
  boltz               reference src/utility.f90:469-498
  manage_fragments    reference src/write_fragments.f90:12-457 (EI and CID record layout :402-441)

The single points run through the batched CUDA entry point (qcxms_b200_egrad_batch); everything else is integer /
small floating-point book-keeping whose results must match the reference digit by digit (record format, rounding).

Deviation that cannot be restated: the reference evaluates `ip_diff(i,1) = fragip(i,1) - fragip(i,0)` right after setting
`fragip(i,1) = 0` (write_fragments.f90:117-120) -- an out-of-bounds read whose value is undefined.  We implement the
evident intent, ip_diff(i,1) = IP_1(i), ip_diff(i,j) = IP_j(i) - IP_(j-1)(i).  Only mchrg = +/-1 is supported here.
"""
import numpy as np

from . import api
from .api import AUTOEV, KB

# src/dftd4.f90:38-72: chemical hardness and third-order parameters used by eself (H..Ar)
GAM = [0.0, 0.47259288, 0.92203391, 0.17452888, 0.25700733, 0.33949086, 0.42195412, 0.50438193, 0.58691863, 0.66931351,
       0.75191607, 0.17964105, 0.22157276, 0.26348578, 0.30539645, 0.34734014, 0.38924725, 0.43115670, 0.47308269]
GAM3 = [0.0, -0.02448, 0.178614, 0.194034, 0.154068, 0.173892, 0.167160, 0.156306, 0.161466, 0.163314, 0.170862, 0.256128,
        0.189060, 0.146310, 0.136686, 0.123558, 0.122070, 0.119424, 0.115368]


_RAD_AA = None


def radii_bohr(num):
    """Covalent radii Rad (reference src/covalent_radii.f90:89-115, bohr) from the same generated table the CUDA side compiles in."""
    global _RAD_AA
    if _RAD_AA is None:
        import os, re
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "params", "elem_tables.h")) as f:
            m = re.search(r"QCXMS_RAD_AA\[[^\]]*\]\s*=\s*\{([^}]*)\}", f.read())
        _RAD_AA = np.array([float(v) for v in m.group(1).split(",")])
    return _RAD_AA[np.asarray(num, dtype=np.int64)] * (1.0 / 0.52917726)


def get_core_e(z):
    """reference src/iniqm.f90:17-41 (H..Ar)"""
    return 0 if z <= 2 else (2 if z <= 10 else 10)


def electrons_amount(num, chrg):
    """reference src/iniqm.f90:45-63 -> (nel, nb, z)"""
    z = np.array([int(a) - get_core_e(int(a)) for a in num], dtype=np.float64)
    nel = int(z.sum()) - int(chrg)
    return nel, nel // 2, z


def getspin(num, chrg):
    """reference src/utility.f90:449-464"""
    j = int(np.sum(num)) - abs(int(chrg))
    return -1 if j < 1 else 1 + j % 2


def eself(num, z):
    """molecular/atomic self energy of a system without electrons, reference src/iniqm.f90:430-451"""
    e = 0.0
    for a, zi in zip(num, z):
        e = e + 0.5 * zi ** 2 * GAM[int(a)] + zi ** 3 * GAM3[int(a)] / 3.0
    return e


def avg_frag_struc(num, axyz, lst, nfrag):
    """Compact per-fragment atom lists (reference src/analyse.f90:452-502): returns natf, iatf, xyzf (lists over fragments)."""
    num = np.asarray(num); axyz = np.asarray(axyz, dtype=np.float64).reshape(-1, 3); lst = np.asarray(lst)
    natf, iatf, xyzf = [], [], []
    for f in range(1, nfrag + 1):
        idx = np.nonzero(lst == f)[0]
        natf.append(len(idx)); iatf.append(num[idx].astype(np.int32)); xyzf.append(axyz[idx].copy())
    return natf, iatf, xyzf


def _gpu_energies(jobs, etemp):
    """jobs: list of (num, xyz, charge).  One batched launch per (composition, charge) group; returns energies, stats."""
    out_e, out_s = [0.0] * len(jobs), [0] * len(jobs)
    groups = {}
    for k, (num, xyz, chrg) in enumerate(jobs):
        groups.setdefault((tuple(int(a) for a in num), int(chrg)), []).append(k)
    for (key, chrg), ks in groups.items():
        num = np.array(key, dtype=np.int32)
        xyz = np.stack([np.asarray(jobs[k][1], dtype=np.float64) for k in ks])
        res = api.egrad_batch(num, xyz, chrg, getspin(num, chrg), api.gfn2_xtb, etemp)
        for j, k in enumerate(ks):
            out_e[k], out_s[k] = float(res["energy"][j]), int(res["stat"][j])
    return out_e, out_s


def eqm_energies(jobs, etemp, energies=_gpu_energies):
    """eqm() for a list of (num, xyz, charge) with the xtb2 branch (reference src/iniqm.f90:397-413): systems without
    electrons get eself; a failed single point is fatal in the reference (`error stop`), here a RuntimeError."""
    e = [0.0] * len(jobs)
    todo = []
    for k, (num, xyz, chrg) in enumerate(jobs):
        nel, _, z = electrons_amount(num, chrg)
        if nel == 0:
            e[k] = eself(num, z)
        else:
            todo.append(k)
    if todo:
        es, st = energies([jobs[k] for k in todo], etemp)
        for k, ek, sk in zip(todo, es, st):
            if sk != 0:
                raise RuntimeError("[Fatal] Calculation in tblite library failed")      # src/iniqm.f90:411-413
            e[k] = ek
    return e


def analyse(num, axyz, lst, nfrag, mchrg=1, energies=_gpu_energies):
    """Fragment IPs (eV) of the average fragment structures (reference src/analyse.f90:14-448, iprog = 8).
    Returns dict(fragip[nfrag], natf, iatf, xyzf, ipok, e_neut, e_ion, rf) -- rf: inter-fragment distances of the
    atomic-number-weighted centres (Angstrom).  The reference retries a failed IP set with the same method and then falls
    back to GFN1 / ORCA (useprog :218-223), which are out of scope here: after two tries the IPs are zeroed, ipok = False."""
    if abs(mchrg) != 1:
        raise NotImplementedError("only singly charged ions")
    natf, iatf, xyzf = avg_frag_struc(num, axyz, lst, nfrag)
    cema = np.array([(x * a[:, None]).sum(0) / a.sum() for a, x in zip(iatf, xyzf)])
    rf = np.array([[np.linalg.norm(cema[i] - cema[j]) * 0.52917726 for j in range(nfrag)] for i in range(nfrag)])
    out = dict(natf=natf, iatf=iatf, xyzf=xyzf, rf=rf, fragip=np.zeros(nfrag), ipok=True, e_neut=np.zeros(nfrag), e_ion=np.zeros(nfrag))
    if nfrag <= 1:
        return out
    etemp = 300.0
    for itry in (1, 2):
        jobs = [(iatf[i], xyzf[i], 0) for i in range(nfrag)] + [(iatf[i], xyzf[i], mchrg) for i in range(nfrag)]
        e = eqm_energies(jobs, etemp, energies)
        ipok = True
        fragip = np.zeros(nfrag)
        for i in range(nfrag):
            e_neut, e_ion = e[i], e[nfrag + i]
            if abs(e_neut) < 1e-10 or abs(e_ion) < 1e-10:      # 'QM code failure' (src/iniqm.f90:418-424)
                ipok = False
            if e_ion != 0 and e_neut != 0:
                fragip[i] = (e_ion - e_neut) * AUTOEV
                if mchrg < 0 and (fragip[i] > 40.0 or fragip[i] < -35.0):
                    ipok = False
                if mchrg == 1 and (fragip[i] < 0.0 or fragip[i] > 50.0):
                    ipok = False
        out.update(fragip=fragip, ipok=ipok, e_neut=np.array(e[:nfrag]), e_ion=np.array(e[nfrag:]))
        if ipok:
            return out
    out.update(fragip=np.zeros(nfrag), ipok=False)
    return out


def boltz(units, temp, ip):
    """Boltzmann populations of the charge over fragments (reference src/utility.f90:469-498); ip in eV (units 2)."""
    const = {1: 627.50947428, 2: AUTOEV}[units]
    f = temp * KB * const
    ip = np.asarray(ip, dtype=np.float64)
    w = np.exp(-ip / f)
    return w / w.sum()


def _fortran_int(v, w):
    s = "%d" % v
    return "*" * w if len(s) > w else s.rjust(w)


def res_line(charge, mchrg, itrj, isec, ifrag, pairs, icoll=None):
    """One qcxms.res / qcxms_cid.res record, format '(F10.7,i3,2i5,2i2,2x,i3,2x,20(i4,i3))' (reference
    src/write_fragments.f90:402-441).  pairs = [(Z or 100 + isotope mass, count), ...].  EI passes one integer fewer
    than the edit descriptors expect, so its columns are shifted exactly as in the reference."""
    items = [int(mchrg), int(itrj)] + ([int(icoll)] if icoll is not None else []) + [int(isec), int(ifrag), len(pairs)]
    for z, c in pairs:
        items += [int(z), int(c)]
    desc = [3, 5, 5, 2, 2, "x2", 3, "x2"] + [4, 3] * 20
    s = "%10.7f" % charge
    if len(s) > 10:
        s = "*" * 10
    pad = ""
    it = iter(items)
    nxt = next(it, None)
    for d in desc:
        if nxt is None:
            break
        if isinstance(d, str):
            pad += " " * int(d[1:])
            continue
        s += pad + _fortran_int(nxt, d)
        pad = ""
        nxt = next(it, None)
    return s


def fragat_pairs(num, lst, ifrag, imass=None):
    """(type, count) pairs of fragment ifrag in ascending type order, as manage_fragments reads them out of fragat(200,10)
    (reference src/fragments.f90:28-40, src/write_fragments.f90:383-390); type = Z or 100 + isotope mass."""
    cnt = {}
    for k in range(len(num)):
        if lst[k] != ifrag:
            continue
        j = int(num[k])
        if imass is not None and imass[k] > 0:
            j = 100 + int(imass[k])
        cnt[j] = cnt.get(j, 0) + 1
    return sorted(cnt.items())


def manage_fragments(num, mass, axyz, lst, qat, aTlast, itrj, isec, mchrg=1, chrgcont=1.0, btf=1.0, maxsec=7, icoll=None,
                     imass=None, energies=_gpu_energies):
    """Charge assignment and records for one finished md() call (reference src/write_fragments.f90:12-457), EI / CID with
    |mchrg| = 1.  `lst` is the fragment list md() returned (the reference recomputes it from the final coordinates with
    fragment_structure, :78).  Returns dict(nfrag, nfrag_ok, lines (written to qcxms.res now), asave (held back for the
    fragment that continues), tcont (1-based, 0: none), chrgcont, mchrg, fragchrg3, fragip, fragm, ipok)."""
    num = np.asarray(num); lst = np.asarray(lst); mass = np.asarray(mass, dtype=np.float64)
    # fragmass: number of fragments with mass, masses in amu (reference src/fragments.f90:10-84)
    fragm = [mass[lst == f].sum() * (1.0 / 1.660539040e-27) * 9.10938356e-31 for f in range(1, 11)]
    nfrag = sum(1 for m in fragm if m > 0)
    out = dict(nfrag=nfrag, nfrag_ok=nfrag <= 5, lines=[], asave=None, tcont=0, chrgcont=chrgcont, mchrg=mchrg, fragchrg3=None,
               fragip=None, fragm=fragm[:nfrag], ipok=True)
    if nfrag > 5:
        return out
    nfrag = int(lst.max())
    an = analyse(num, axyz, lst, nfrag, mchrg, energies)
    natf, iatf = an["natf"], an["iatf"]
    fragchrg3 = np.zeros(nfrag)
    if nfrag > 1:
        ip_diff = an["fragip"].copy()
        w = boltz(2, aTlast * btf, ip_diff)
        if mchrg < 0:
            w = -w
        fragchrg3 = w.copy()
    else:
        fragchrg3[:] = chrgcont
    tcont = 0
    if mchrg > 0:
        largest = -1.0
        for i in range(nfrag):
            if float(natf[i]) * fragchrg3[i] > largest:
                largest = float(natf[i]) * fragchrg3[i]; tcont = i + 1
    else:
        largest = 1.0
        for i in range(nfrag):
            if float(natf[i]) * fragchrg3[i] < largest:
                largest = float(natf[i]) * fragchrg3[i]; tcont = i + 1
    if nfrag == 1 or isec == maxsec + 1:
        tcont = 0
    else:
        fragchrg3 = fragchrg3 * abs(chrgcont) / abs(mchrg)
        chrgcont = fragchrg3[tcont - 1]
        nearest = int(np.floor(abs(chrgcont) + 0.5) * np.sign(chrgcont))      # nint
        if nearest > 0 and mchrg > 0: mchrg = nearest
        if nearest == 0 and mchrg > 0: mchrg = 1
        if nearest < 0 and mchrg < 0: mchrg = nearest
        if nearest == 0 and mchrg < 0: mchrg = -1
    lines, asave = [], None
    for j in range(1, nfrag + 1):
        rec = res_line(fragchrg3[j - 1], mchrg, itrj, isec, j, fragat_pairs(num, lst, j, imass), icoll=icoll)
        if tcont > 0 and tcont == j:
            asave = rec
        elif tcont == 0 and icoll is not None:
            asave = rec      # CID (method 3): held back as well, each fragment overwriting the last (src/write_fragments.f90:411-421)
        else:
            lines.append(rec)
    out.update(nfrag=nfrag, lines=lines, asave=asave, tcont=tcont, chrgcont=chrgcont, mchrg=mchrg, fragchrg3=fragchrg3,
               fragip=an["fragip"], ipok=an["ipok"], natf=natf, fragq=[float(np.asarray(qat)[lst == f].sum()) for f in range(1, nfrag + 1)])
    return out


def spectrum_from_records(lines, nbins=512):
    """Charge-weighted stick spectrum (nominal masses from the most abundant isotopes) from qcxms.res records -- the
    quantity PlotMS accumulates before the isotope-pattern expansion (PlotMS itself is not part of the reference tree)."""
    nominal = {1: 1, 2: 4, 6: 12, 7: 14, 8: 16, 9: 19, 16: 32, 17: 35, 18: 40}
    bins = np.zeros(nbins)
    for ln in lines:
        chg = float(ln[:10])
        rest = ln[10:].split()
        ntypes = None
        # EI: mchrg itrj isec ifrag ntypes Z1 c1 ...; CID has icoll after itrj.  ntypes is the field followed by 2 * ntypes values.
        vals = [int(v) for v in rest]
        for pos in (4, 5):
            if pos < len(vals) and len(vals) - pos - 1 == 2 * vals[pos]:
                ntypes = pos
        if ntypes is None:
            continue
        m = 0
        for k in range(vals[ntypes]):
            z, c = vals[ntypes + 1 + 2 * k], vals[ntypes + 2 + 2 * k]
            m += (z - 100 if z > 100 else nominal[z]) * c
        if 0 <= m < nbins:
            bins[m] += abs(chg)
    return bins
