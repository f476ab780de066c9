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

`run_cid` is the same for a CID run: the collision loop around cid() with the mean-free-path MDs between two collisions

  main.F90:1490-1592   number of collisions of a trajectory (fullauto / collauto / collno / collsec / maxcoll run types)
  main.F90:1600-1770   cidlp: direc from the drift of the centre of mass, cid(), manage_fragments, cut out the charged fragment
  main.F90:1772-1975   MFPloop: md() with method 3 and icoll >= 1, nmax = 100 nuc scaled by isec, fragment cut-out, repeat while it fragments
  main.F90:1981-2128   stop rules: small fragment, low mass, slow ion / low E(COM), number of fragmentations, number of collisions
  main.F90:2133-2163   the held-back record of the charged fragment is written when a trajectory ends early
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import fragments as fr

# Groups of one generation (different compositions) are independent launches: they are driven from a few host threads so that
# their kernels, each of which fills only part of the GPU once the ensemble has fragmented into many compositions, overlap
# (every ensemble / collision batch has its own stream; ctypes releases the GIL during the calls).
GROUP_THREADS = 8


def _map_groups(fn, items):
    items = list(items)
    if len(items) <= 1 or GROUP_THREADS <= 1:
        return [fn(it) for it in items]
    with ThreadPoolExecutor(max_workers=min(GROUP_THREADS, len(items))) as ex:
        return list(ex.map(fn, items))


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
        def run_group(item):
            (_, mc, nm, nfe, isec), members = item
            g0 = members[0]
            return md_batch(g0["num"], g0["mass"], [s["xyz"] for s in members], [s["velo"] for s in members], [s["velof"] for s in members],
                            [s["eimp"] for s in members], [s["tadd"] for s in members], mc, nm, nfe, isec, tstep_fs, etemp)
        for (_, members), out in zip(groups.items(), _map_groups(run_group, groups.items())):
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
    try:
        out = fr.manage_fragments(s["num"], s["mass"], r["axyz"], r["list"], r["achrg"], aTlast=float(r["aTlast"]), itrj=s["itrj"], isec=s["isec"],
                                  mchrg=s["mchrg"], chrgcont=s["chrgcont"], btf=btf, maxsec=maxsec, energies=energies)
    except RuntimeError as err:                   # the reference process of this trajectory `error stop`s (src/iniqm.f90:411-413)
        gen.update(fatal=str(err))
        s["alive"] = False
        return
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


# ------------------------------------------------------------------------------------------------------------------- CID
AATOAU = 1.0 / 0.52917726                 # reference src/xtb_mctc_convert.f90 (aatoau = 1 / autoaa)
MSTOAU = 1.0 / 2.18769126364e+06
KB = 3.166808578545117e-06
FSTOAU = 41.3413733365614
AUTOEV = 27.21138505
AMUTOAU = 1.660539040e-27 * (1.0 / 9.10938356e-31)


def collision_setup(num, xyz, gas="ar", tgas=300.0, pgas=0.132, lchamb=0.2):
    """Molecular radius, cross section, mean free path and expected number of collisions (reference src/cid.f90:1163-1202).
    Returns (r_mol / m, cross / m^2, mfpath / m, calc_collisions)."""
    from .api import GASES
    rad = fr.radii_bohr(num)
    xyz = np.asarray(xyz, dtype=np.float64)
    cg = xyz.sum(0) / len(xyz)
    rtot = 0.0
    for i in range(len(xyz)):
        mol_rad = np.sqrt(((xyz[i] - cg) ** 2).sum()) + rad[i]
        if mol_rad > rtot:
            rtot = mol_rad
    r_atom = (GASES[gas.lower()][2] / AATOAU) * 1e-10
    r_mol = (rtot / AATOAU) * 1e-10
    cross = np.pi * ((r_mol + r_atom) ** 2)
    mfpath = (1.38064852e-23 * tgas) / (cross * pgas)
    return r_mol, cross, mfpath, lchamb / mfpath


def vary_collisions(calc_collisions, dum, dum2):
    """Box-Muller variation of the number of collisions (reference src/boxmuller.f90:19-44); dum, dum2: the two uniform random numbers."""
    sigma = calc_collisions * 0.12
    z0 = np.sqrt(-2.0 * np.log(dum)) * np.cos(2.0 * np.pi * dum2)
    n = int(np.floor(z0 * sigma + calc_collisions + 0.5)) if z0 * sigma + calc_collisions >= 0 else -int(np.floor(-(z0 * sigma + calc_collisions) + 0.5))
    return max(n, 0)


def gpu_cid_batch(cfg, num, mass, icoll, xyz, velo, rnd, velo_cm, direc, collided):
    """cid() for a group of ions of one composition on the GPU (qcxms_b200_cid_batch); list of per-trajectory result dicts."""
    from . import api
    out = api.cid(cfg, num, mass, icoll, np.asarray(xyz), np.asarray(velo), np.asarray(rnd), velo_cm=None if icoll == 1 else np.asarray(velo_cm),
                  direc=None if icoll == 1 else np.asarray(direc), collided=np.asarray(collided, dtype=np.int32))
    return [{k: v[i] for k, v in out.items()} for i in range(len(xyz))]


def gpu_mfp_batch(num, mass, xyz, velo, new_velo, icoll, isec, mchrg, nmax, tstep_fs, etemp):
    """mean-free-path md() for a group of ions of one composition on the GPU (ensemble in set_mfp mode)."""
    from . import api
    nt = len(xyz)
    ens = api.Ensemble(num, mass, nt, mchrg=mchrg, tstep_fs=tstep_fs, nmax=nmax, nfragexit=3, exit_rules=True, etemp=etemp, isec=isec)
    try:
        z = np.zeros(nt)
        ens.set_all(np.asarray(xyz), np.asarray(velo), np.ones((nt, len(num))), z, z)
        ens.set_mfp(icoll, np.asarray(new_velo, dtype=np.float64))
        ens.run_md()
        res = ens.results()
        res["new_velo"] = ens.new_velo()
    finally:
        ens.close()
    return [{k: (v[i] if isinstance(v, np.ndarray) else v) for k, v in res.items()} for i in range(nt)]


def gpu_esi_batch(num, mass, xyz, velo, tscale, e_scale, pretadd, mchrg, nmax, tstep_fs, etemp):
    """heating MD before the first collision (md() with method 3, icoll 0, starting_md) for ions of one composition and one length"""
    from . import api
    nt = len(xyz)
    ens = api.Ensemble(num, mass, nt, mchrg=mchrg, tstep_fs=tstep_fs, nmax=nmax, nfragexit=3, exit_rules=True, etemp=etemp, isec=1)
    try:
        ens.set_all(np.asarray(xyz), np.asarray(velo), np.ones((nt, len(num))), np.full(nt, float(e_scale)), np.full(nt, float(pretadd)))
        ens.set_esi(tscale)
        ens.run_md()
        res = ens.results()
    finally:
        ens.close()
    return [{k: (v[i] if isinstance(v, np.ndarray) else v) for k, v in res.items()} for i in range(nt)]


def _cut_fragment(s, xyz, velo, lst, tcont):
    """Atoms of fragment tcont (all atoms for tcont == 0), recentred on sum(Z r) / sum(Z) (sic) -- main.F90:1672-1754, 1911-1963.
    Returns False when the rest is too small / too light to continue."""
    sel = np.nonzero(np.asarray(lst) == tcont)[0] if tcont > 0 else np.arange(len(s["num"]))
    if tcont > 0:
        s["frag_counter"] += 1
    iatn = s["num"][sel]
    xyzn = np.asarray(xyz, dtype=np.float64)[sel]
    cema = (xyzn * iatn[:, None]).sum(0) / iatn.sum()
    parent_mass = s["mass"]
    s["num"] = iatn.astype(np.int32)
    s["mass"] = parent_mass[sel]
    if len(sel) <= 7:
        s["small"] = True
        return False
    # sic: the reference sums the first nuc entries of the PARENT's mass array (it is re-filled with the fragment's masses only
    # after this test, main.F90:1727, 1936), not the masses of the new fragment
    if parent_mass[:len(sel)].sum() / AMUTOAU <= s["minmass"]:
        s["littlemass"] = True
        return False
    s["xyz"] = xyzn - cema
    s["velo"] = np.asarray(velo, dtype=np.float64)[sel]
    return True


def run_cid(num, mass, xyz, velo, mchrg=1, gas="ar", elab=40.0, ecom=0.0, eexact=False, manual_dist=0, tstep_fs=0.5, etemp=-1.0, btf=1.0,
            maxsec=7, run_type="fullauto", set_coll=10, max_coll=0, collno=(0, 0, 0), collsec=(0, 0, 0), tgas=300.0, pgas=0.132, lchamb=0.2,
            minmass=45, first_itrj=1, seed=0, cid_ntot=15000, mfp_nmax=None, cid_batch=gpu_cid_batch, mfp_batch=gpu_mfp_batch,
            energies=fr._gpu_energies, esi_ev=0.0, esi_nmax=None, esi_batch=gpu_esi_batch):
    """CID production run for ntraj ions given as arrays with a leading [ntraj] axis.  esi_ev = 0: the reference's `noesi`; esi_ev > 0
    (keyword `esi <eV>`): every ion whose internal energy is below esi_ev is first heated to it by the thermostatted MD of
    main.F90:1243-1362 (md() with method 3, icoll 0, starting_md; nmax = nint(2.5 (T_target - T)), pretadd = 3/4 of its length).  An ion that
    fragments while being heated goes on with its charged fragment straight to the collisions (the reference runs further heating MDs
    with isec > 1 first: not reproduced).  run_type: "fullauto" | "collauto" | "maxcoll" | "collno" | "collsec" (main.F90:1490-1592).
    Random numbers (9 per cid() call, 2 per vary_collisions, 1 per collauto redraw) come from one numpy Generator per trajectory seeded
    with (seed, itrj), so a trajectory does not depend on how the ensemble is batched.  mfp_nmax / cid_ntot shorten the runs (tests).
    Returns dict(records = qcxms_cid.res lines, per_traj = [dict(itrj, events, records)])."""
    from . import api
    num0 = np.asarray(num, dtype=np.int32)
    nt = len(xyz)
    # sic: the mass of ONE gas atom also for N2 (gas%mIatom = 14.007, input.f90 "IATOM N2"; beta in cid.f90:347 and the E_COM stop
    # rule of main.F90 use it as it is) -- the device side (qx_cid.cuh) does the same
    gas_mass = api.GASES[gas.lower()][1] * AMUTOAU
    trj = []
    for t in range(nt):
        itrj = first_itrj + t
        s = dict(itrj=itrj, rng=np.random.default_rng([int(seed), itrj]), num=num0.copy(), mass=np.asarray(mass, dtype=np.float64).copy(),
                 xyz=np.array(xyz[t], dtype=np.float64), velo=np.array(velo[t], dtype=np.float64), mchrg=int(mchrg), chrgcont=float(mchrg),
                 icoll=0, isec=1, frag_counter=0, save_counter=0, new_counter=0, collisions=0, new_velo=0.0, direc=np.zeros(3), collided=0,
                 cm1=np.zeros(3), cm2=np.zeros(3), asave=None, small=False, littlemass=False, minmass=minmass, phase="cid", tcont=0,
                 records=[], events=[])
        if run_type == "fullauto":
            s["collisions"] = vary_collisions(collision_setup(s["num"], s["xyz"], gas, tgas, pgas, lchamb)[3], *s["rng"].random(2))
        elif run_type == "collauto":
            s["collisions"] = int(set_coll)
        elif run_type == "collsec":
            s["collisions"] = int(set_coll)
            s["new_counter"] = collsec[2] if itrj % 20 == 0 else (collsec[1] if itrj % 3 == 0 else collsec[0])
        elif run_type == "collno":
            s["collisions"] = collno[2] if itrj % 10 == 0 else (collno[1] if itrj % 3 == 0 else collno[0])
        elif run_type == "maxcoll":
            s["collisions"], s["new_counter"] = int(max_coll), 1
        else:
            raise ValueError("unknown CID run type " + str(run_type))
        if len(s["num"]) <= 7:          # "Simulation stopped - too small molecule for collisions" (main.F90:1483)
            s["phase"] = "done"
        trj.append(s)

    def finish(s, write_asave):
        if write_asave and s["asave"] is not None:
            s["records"].append(s["asave"])
        s["phase"] = "done"

    def after_fragments(s, out, xyz_, velo_, lst):
        """manage_fragments bookkeeping + fragment cut-out; returns False when the trajectory ended"""
        s["records"] += out["lines"]
        s["asave"], s["mchrg"], s["chrgcont"], s["tcont"] = out["asave"], out["mchrg"], out["chrgcont"], out["tcont"]
        if not _cut_fragment(s, xyz_, velo_, lst, out["tcont"]):
            finish(s, True)                      # small / littlemass: main.F90:2133-2155
            return False
        return True

    # ---- ESI: heat the ions to the requested internal energy before the first collision (main.F90:1243-1362)
    if esi_ev > 0.0:
        groups = {}
        for s in trj:
            if s["phase"] != "cid":
                continue
            nuc = len(s["num"])
            ekin = 0.5 * (s["mass"][:, None] * s["velo"] ** 2).sum()
            temp = ekin / (0.5 * 3 * nuc * KB)
            ene1 = temp * (0.5 * 3 * nuc * KB) * AUTOEV
            e_scale = esi_ev if esi_ev - ene1 > 0 else 0.0
            if e_scale <= 0:
                continue                                          # "! No Scaling !"
            tscale = (e_scale * 2.0 / 3.0) / (nuc * KB * AUTOEV)
            nmax = int(np.floor((tscale - temp) * 2.5 + 0.5)) if not esi_nmax else int(esi_nmax)
            pretadd = float(nmax) * FSTOAU * tstep_fs * 0.75
            groups.setdefault((tuple(int(a) for a in s["num"]), s["mchrg"], nmax), []).append((s, tscale, e_scale / AUTOEV, pretadd))

        def run_esi(item):
            (_, mc, nmax), mem = item
            g0 = mem[0][0]
            return esi_batch(g0["num"], g0["mass"], [m[0]["xyz"] for m in mem], [m[0]["velo"] for m in mem], mem[0][1], mem[0][2], mem[0][3], mc, nmax,
                             tstep_fs, etemp)
        for ((_, mc, nmax), mem), out in zip(groups.items(), _map_groups(run_esi, groups.items())):
            for (s, tscale, _, _), r in zip(mem, out):
                md_ok = bool(r["mdok"]) and int(r["status"]) == 1
                s["events"].append(dict(kind="esi", icoll=0, isec=1, nat=len(s["num"]), nstep=int(r["nstep"]), nfrag=int(r["nfrag"]), md_ok=md_ok,
                                        tscale=float(tscale)))
                if not md_ok:
                    finish(s, False)                 # the reference process stops here (main.F90:1398)
                    continue
                try:
                    mf = fr.manage_fragments(s["num"], s["mass"], r["axyz"], r["list"], r["achrg"], aTlast=float(r["aTlast"]), itrj=s["itrj"],
                                             isec=1, mchrg=s["mchrg"], chrgcont=s["chrgcont"], btf=btf, maxsec=maxsec, icoll=0, energies=energies)
                except RuntimeError as err:
                    s["events"].append(dict(kind="fatal", icoll=0, nstep=0, nfrag=0, msg=str(err)))
                    finish(s, False)
                    continue
                if not mf["nfrag_ok"]:
                    finish(s, False)
                    continue
                after_fragments(s, mf, r["xyz"], r["velo"], r["list"])

    while any(s["phase"] != "done" for s in trj):
        # ---- collisions
        groups = {}
        for s in trj:
            if s["phase"] == "cid":
                s["isec"] = 1
                s["icoll"] += 1
                if s["icoll"] != 1:
                    d = s["cm2"] - s["cm1"]
                    s["direc"] = d / np.sqrt((d * d).sum())
                groups.setdefault((tuple(int(a) for a in s["num"]), s["mchrg"], min(s["icoll"], 2)), []).append(s)
        # trajectories of one composition and one collision index run as one batch
        batches = {}
        for (comp, mc, _), members in groups.items():
            for s in members:
                batches.setdefault((comp, mc, s["icoll"]), []).append(s)

        def run_collision(item):
            (_, mc, icoll), mem = item
            g0 = mem[0]
            cfg = api.cid_config(mchrg=mc, gas=gas, tstep_fs=tstep_fs, elab=elab, ecom=ecom, eexact=eexact, manual_dist=manual_dist,
                                 ntot=cid_ntot, etemp=max(etemp, 0.0))
            rnd = [s["rng"].random(9) for s in mem]
            return cid_batch(cfg, g0["num"], g0["mass"], icoll, [s["xyz"] for s in mem], [s["velo"] for s in mem], rnd,
                             [s["new_velo"] for s in mem], [s["direc"] for s in mem], [s["collided"] for s in mem])
        for ((_, mc, icoll), mem), out in zip(batches.items(), _map_groups(run_collision, batches.items())):
            if True:
                for s, r in zip(mem, out):
                    s["collided"] = int(r["collided"])
                    s["events"].append(dict(kind="cid", icoll=icoll, nat=len(s["num"]), nstep=int(r["nstep"]), nfrag=int(r["nfrag"]),
                                            stopcid=int(r["stopcid"]), velo_cm=float(r["velo_cm"])))
                    if int(r["stopcid"]) or int(r["status"]) != 1:
                        finish(s, True)              # "run aborted, last structure saved" (main.F90:2157-2163)
                        continue
                    s["new_velo"] = float(r["velo_cm"])
                    if icoll == 1:
                        s["direc"] = np.array(r["direc"], dtype=np.float64)
                    try:
                        mf = fr.manage_fragments(s["num"], s["mass"], r["axyz"], r["list"], r["achrg"], aTlast=float(r["aTlast"]), itrj=s["itrj"],
                                                 isec=1, mchrg=s["mchrg"], chrgcont=s["chrgcont"], btf=btf, maxsec=maxsec, icoll=icoll,
                                                 energies=energies)
                    except RuntimeError as err:      # the reference process of this trajectory `error stop`s (src/iniqm.f90:411-413)
                        s["events"].append(dict(kind="fatal", icoll=icoll, nstep=0, nfrag=0, msg=str(err)))
                        finish(s, False)
                        continue
                    if not mf["nfrag_ok"]:
                        finish(s, False)
                        continue
                    if after_fragments(s, mf, r["xyz"], r["velo"], r["list"]):
                        s["phase"] = "mfp"
        # ---- mean-free-path MDs, repeated while they fragment
        groups = {}
        for s in trj:
            if s["phase"] == "mfp":
                s["isec"] += 1
                mtot = s["mass"].sum()
                s["cm1"] = (s["xyz"] * s["mass"][:, None]).sum(0) / mtot
                nmax = len(s["num"]) * 100
                if s["isec"] == 3: nmax = int(nmax * 0.75)
                if s["isec"] == 4: nmax = int(nmax * 0.6)
                if s["isec"] >= 5: nmax = int(nmax * 0.5)
                nmax = min(max(nmax, 1000), 10000)
                if mfp_nmax:
                    nmax = int(mfp_nmax)
                groups.setdefault((tuple(int(a) for a in s["num"]), s["mchrg"], nmax, s["icoll"], s["isec"]), []).append(s)
        def run_mfp(item):
            (_, mc, nmax, icoll, isec), mem = item
            g0 = mem[0]
            return mfp_batch(g0["num"], g0["mass"], [s["xyz"] for s in mem], [s["velo"] for s in mem], [s["new_velo"] for s in mem], icoll, isec,
                             mc, nmax, tstep_fs, etemp)
        for ((_, mc, nmax, icoll, isec), mem), out in zip(groups.items(), _map_groups(run_mfp, groups.items())):
            for s, r in zip(mem, out):
                md_ok = bool(r["mdok"]) and int(r["status"]) == 1
                s["events"].append(dict(kind="mfp", icoll=icoll, isec=isec, nat=len(s["num"]), nstep=int(r["nstep"]), nfrag=int(r["nfrag"]),
                                        md_ok=md_ok, new_velo=float(r["new_velo"])))
                s["new_velo"] = float(r["new_velo"])
                mtot = s["mass"].sum()
                s["cm2"] = (np.asarray(r["xyz"]) * s["mass"][:, None]).sum(0) / mtot
                if not md_ok:
                    finish(s, False)                 # "the run is just not further counted" (main.F90:1896-1903)
                    continue
                try:
                    mf = fr.manage_fragments(s["num"], s["mass"], r["axyz"], r["list"], r["achrg"], aTlast=float(r["aTlast"]), itrj=s["itrj"],
                                             isec=isec, mchrg=s["mchrg"], chrgcont=s["chrgcont"], btf=btf, maxsec=maxsec, icoll=icoll,
                                             energies=energies)
                except RuntimeError as err:
                    s["events"].append(dict(kind="fatal", icoll=icoll, nstep=0, nfrag=0, msg=str(err)))
                    finish(s, False)
                    continue
                if not mf["nfrag_ok"]:
                    finish(s, False)
                    continue
                if not after_fragments(s, mf, r["xyz"], r["velo"], r["list"]):
                    continue
                if mf["tcont"] > 0:
                    continue                          # fragmented: another mean-free-path MD of the charged fragment
                # ---- stop rules before the next collision (main.F90:1981-2128)
                summass = s["mass"].sum()
                e_kin = 0.5 * summass * ((s["new_velo"] * MSTOAU) ** 2)
                e_com = (gas_mass / (gas_mass + summass)) * e_kin * AUTOEV
                if s["new_velo"] <= 800 or e_com <= 0.85:
                    finish(s, True)
                    continue
                if run_type in ("collsec", "collno", "maxcoll") and s["new_counter"] > 0 and s["frag_counter"] >= s["new_counter"]:
                    finish(s, True)
                    continue
                if run_type == "fullauto" and s["frag_counter"] > s["save_counter"]:
                    s["save_counter"] = s["frag_counter"]
                    s["collisions"] = vary_collisions(collision_setup(s["num"], s["xyz"], gas, tgas, pgas, lchamb)[3], *s["rng"].random(2))
                    if s["collisions"] == 0 and s["icoll"] != 1:
                        finish(s, True)
                        continue
                if run_type == "collauto" and s["frag_counter"] > s["save_counter"]:      # (sic) save_counter is not advanced here
                    if s["collisions"] > 0:
                        dep = int(np.floor(len(s["num"]) / 10.0 + 0.5))
                        s["collisions"] = s["icoll"] + int(np.floor((dep + 1) * s["rng"].random()))
                    elif s["icoll"] != 1:
                        finish(s, True)
                        continue
                if s["icoll"] >= s["collisions"]:
                    if s["asave"] is not None:
                        s["records"].append(s["asave"])
                    s["phase"] = "done"
                    continue
                s["phase"] = "cid"
    records = []
    for s in trj:
        records += s["records"]
    return dict(records=records, per_traj=[dict(itrj=s["itrj"], events=s["events"], records=s["records"]) for s in trj])
