"""EI production cascade around md(): primary trajectory -> fragment bookkeeping -> secondary trajectories of the fragment
that keeps the charge, batched over the whole ensemble.  Host-side mirror of the reference's production loop

  main.F90:2199-2370   loop: isec += 1; iniqm; md(); manage_fragments(); cut out fragment `tcont`; cycle
  main.F90:2252        nfragexit = 2 for isec > 1
  main.F90:2290-2296   nmax <- max(nmax - nmax0/5, nmax0/2); fragstate == 2 -> nmax0/3; nuc <= 10 -> nmax0/10; nuc <= 3 -> stop
  main.F90:2298-2345   new system: atoms of fragment tcont, recentred on sum(Z r)/sum(Z), velocities kept, velof = tadd = eimp = 0

The reference runs this loop once per trajectory directory (bin/pqcxms); here every generation of the cascade is one set of
ensemble launches: trajectories that are still alive are grouped by (composition, charge, step budget) and each group runs as
one `Ensemble` (qcxms_b200_ensemble_*), the 2 nfrag single points per trajectory of `analyse` go through the batched egrad.
The MD and single-point back ends are parameters so that the same driver runs on the CPU oracle in the tests.
"""
import numpy as np

from . import fragments as fr


def gpu_md_batch(num, mass, xyz, velo, velof, eimp, tadd, mchrg, nmax, nfragexit, isec, tstep_fs, etemp):
    """md() for a group of trajectories of one composition on the GPU; returns a list of per-trajectory result dicts."""
    from . import api
    nt = len(xyz)
    ens = api.Ensemble(num, mass, nt, mchrg=mchrg, tstep_fs=tstep_fs, nmax=nmax, nfragexit=nfragexit, exit_rules=True, etemp=etemp,
                       isec=isec)
    try:
        ens.set_all(np.asarray(xyz), np.asarray(velo), np.asarray(velof), np.asarray(eimp), np.asarray(tadd))
        ens.run_md()
        res = ens.results()
    finally:
        ens.close()
    return [{k: (v[i] if isinstance(v, np.ndarray) else v) for k, v in res.items()} for i in range(nt)]


def run_ei(num, mass, xyz, velo, velof, eimp, tadd, mchrg=1, nmax=10000, maxsec=7, nfragexit=3, tstep_fs=0.5, etemp=-1.0, btf=1.0,
           first_itrj=1, md_batch=gpu_md_batch, energies=fr._gpu_energies):
    """Runs the EI cascade for ntraj trajectories given as arrays with a leading [ntraj] axis.
    Returns dict(records = qcxms.res lines in trajectory order, per_traj = list of dicts with the generations of each trajectory)."""
    num0 = np.asarray(num, dtype=np.int32)
    nt = len(xyz)
    nmax0 = int(nmax)
    trj = []
    for t in range(nt):
        trj.append(dict(itrj=first_itrj + t, num=num0.copy(), mass=np.asarray(mass, dtype=np.float64).copy(), xyz=np.array(xyz[t], dtype=np.float64),
                        velo=np.array(velo[t], dtype=np.float64), velof=np.array(velof[t], dtype=np.float64), eimp=float(eimp[t]),
                        tadd=float(tadd[t]), mchrg=int(mchrg), chrgcont=float(mchrg), isec=0, nmax=nmax0, alive=True, asave=None,
                        records=[], generations=[]))
    while any(s["alive"] for s in trj):
        groups = {}
        for s in trj:
            if s["alive"]:
                s["isec"] += 1
                nfe = 2 if s["isec"] > 1 else nfragexit                       # main.F90:2252
                key = (tuple(int(a) for a in s["num"]), s["mchrg"], s["nmax"], nfe, s["isec"])
                groups.setdefault(key, []).append(s)
        for (_, mc, nm, nfe, isec), members in groups.items():
            g0 = members[0]
            out = md_batch(g0["num"], g0["mass"], [s["xyz"] for s in members], [s["velo"] for s in members], [s["velof"] for s in members],
                           [s["eimp"] for s in members], [s["tadd"] for s in members], mc, nm, nfe, isec, tstep_fs, etemp)
            for s, r in zip(members, out):
                _after_md(s, r, nmax0, maxsec, btf, energies)
    records = []
    for s in trj:
        records += s["records"]
    return dict(records=records, per_traj=[dict(itrj=s["itrj"], generations=s["generations"], records=s["records"]) for s in trj])


def _after_md(s, r, nmax0, maxsec, btf, energies):
    """Everything the reference does between two md() calls of one trajectory (main.F90:2263-2366)."""
    md_ok = bool(r["mdok"]) and int(r["status"]) == 1
    gen = dict(isec=s["isec"], nat=len(s["num"]), nstep=int(r["nstep"]), nfrag=int(r["nfrag"]), fragstate=int(r["fragstate"]), md_ok=md_ok)
    s["generations"].append(gen)
    if not md_ok:
        if s["asave"] is not None:
            s["records"].append(s["asave"])
        s["alive"] = False
        return
    out = fr.manage_fragments(s["num"], s["mass"], r["axyz"], r["list"], r["achrg"], aTlast=float(r["aTlast"]), itrj=s["itrj"], isec=s["isec"],
                              mchrg=s["mchrg"], chrgcont=s["chrgcont"], btf=btf, maxsec=maxsec, energies=energies)
    gen.update(tcont=out["tcont"], fragip=None if out["fragip"] is None else [float(v) for v in out["fragip"]])
    if not out["nfrag_ok"]:                       # too many fragments: the run is not counted (main.F90:2280)
        s["alive"] = False
        return
    s["records"] += out["lines"]
    s["asave"], s["mchrg"], s["chrgcont"] = out["asave"], out["mchrg"], out["chrgcont"]
    tcont = out["tcont"]
    if not (tcont > 0 and maxsec > 0):
        s["alive"] = False                        # no fragmentation (or last generation): every record has been written
        return
    nmax = s["nmax"] - nmax0 // 5
    nmax = max(nmax, nmax0 // 2)
    if int(r["fragstate"]) == 2:
        nmax = nmax0 // 3
    sel = np.nonzero(np.asarray(r["list"]) == tcont)[0]
    iatn = s["num"][sel]
    xyzn = np.asarray(r["xyz"], dtype=np.float64)[sel]
    cema = (xyzn * iatn[:, None]).sum(0) / iatn.sum()        # atomic-number weighted centre (sic), main.F90:2312-2316
    if len(sel) <= 10:
        nmax = nmax0 // 10
    if len(sel) <= 3:
        s["records"].append(s["asave"])
        s["alive"] = False
        return
    s.update(num=iatn.astype(np.int32), mass=s["mass"][sel], xyz=xyzn - cema, velo=np.asarray(r["velo"], dtype=np.float64)[sel],
             velof=np.zeros(len(sel)), eimp=0.0, tadd=0.0, nmax=nmax)
