"""Ensemble set-up of an EI run: everything the reference does between the neutral input geometry and the
TMPQCXMS/TMP.<n> directories the production runs read.

* ground-state MD on the GPU: md() with it = -1 (equilibration) and it = 0 (sampling), reference src/main.F90:523-567 -- one
  trajectory (or several independent ones) through `Ensemble.set_gs_mode`;
* initial velocities `mdinitu` (src/mdinit.f90:10-52);
* the impact-excess-energy distribution: `getieeab`, `getmaxiee`, `gauss0`, `poiss0` (src/iee.f90:12-135);
* the per-trajectory draw (src/main.F90:684-886): snapshot selection, electron energy, IEE by rejection sampling, `momap`
  (src/utility.f90:23-65), heating factors `velof` from the MO populations (`get_xtb_egrad_spec` = the reference's getspec output),
  heating time from `calctrelax` (src/impact.f90:58-81);
* `write_directories`: start.xyz + qcxms.start per trajectory (src/utility.f90:327-372).

The Fortran intrinsic random_number is replaced by a numpy Generator handed in by the caller: runs are reproducible but not
stream-identical to a reference run (the reference seeds from the clock unless `iseed` is set).
"""
import os

import numpy as np

from . import startfiles
from .api import AMUTOAU, FSTOAU, Ensemble, get_xtb_egrad_spec, gfn2_xtb

KB = 3.166808578545117e-06
AUTOEV = 27.21138505
EVTOAU = 1.0 / AUTOEV


def irand(n, rng):
    """reference src/utility.f90:502-511: integer in 1..n"""
    r = int(n * rng.random() + 1)
    return n if r > n else r


def mdinitu(mass, e_kin_in, rng, velo=None):
    """reference src/mdinit.f90:10-52: every Cartesian component gets +-sqrt(2 eperat / m)"""
    nat = len(mass)
    velo = np.zeros((nat, 3)) if velo is None else np.array(velo, dtype=np.float64)
    eperat = e_kin_in / (3.0 * nat)
    for i in range(nat):
        v = np.sqrt(2 * eperat / mass[i])
        for c in range(3):
            f = -1.0 if rng.random() > 0.5 else 1.0
            velo[i, c] += v * f
    return velo


def gauss0(iee_a, iee_b, ieeel, x):
    return np.exp(-iee_a * (x - ieeel * iee_b) ** 2 / ieeel)


def poiss0(iee_a, iee_b, ieeel, x):
    z, k = iee_b, 1.0 / iee_a
    t2 = k / ieeel
    t8 = np.log(z / k * ieeel / x)
    t14 = np.exp(t2 * x * (1.0 + t8) - 1.0 * z)
    t17 = (t2 * x + 1.0) ** (-0.5)
    return t14 * t17


def getmaxiee(iee_a, iee_b, ieeel, ityp, exc):
    """reference src/iee.f90:50-80 -> (ieemax, pmax, E_avg)"""
    x, pmax, ieemax, e_avg, m = 0.001, -1.0, 0.0, 0.0, 0.0
    while True:
        val = gauss0(iee_a, iee_b, ieeel, x) if ityp == 0 else poiss0(iee_a, iee_b, ieeel, x)
        if val > pmax:
            pmax, ieemax = val, x
        x = x + 0.01
        e_avg = e_avg + val * x
        m = m + val
        if x >= exc:
            break
    return ieemax, pmax, e_avg / m


def getieeab(ieeel, ityp, exc, nbnd, ieeatm):
    """reference src/iee.f90:12-43 -> (iee_a, iee_b)"""
    st, iee_a, iee_b, k = 0.005, 0.0, 0.0, 0
    while True:
        k += 1
        iee_a = min(iee_a + st, np.float64(np.float32(0.3)))      # min(iee_a, 0.3): the literal is single precision
        iee_b = iee_b + st * 7
        _, _, e_avg = getmaxiee(iee_a, iee_b, ieeel, ityp, exc)
        if k > 10000:
            raise RuntimeError("internal error inside getieeab")
        if e_avg / nbnd >= ieeatm:
            return iee_a, iee_b


def momap(ihomo, emo, edum, rng):
    """reference src/utility.f90:23-65: (mo1, mo2) 1-based, mo2 = 0 for an unpaired electron; emo[k-1] = energy of MO k"""
    i1 = i2 = 0
    dmin = np.inf
    for _ in range(5001):
        mo1, mo2 = irand(ihomo, rng), irand(ihomo, rng)
        vmo = irand(ihomo // 2, rng) + ihomo
        dum = emo[mo1 - 1]
        if mo2 > ihomo // 2:
            mo2 = 0
        else:
            dum = dum + emo[mo2 - 1]
        dum = dum + emo[vmo - 1]
        delta = abs(dum - edum)
        if delta < dmin:
            dmin, i1, i2 = delta, mo1, mo2
    return i1, i2


def calctrelax(emo, na, i, trelax):
    """reference src/impact.f90:58-81: sum over the MOs above i (1-based) up to na of trelax exp(alp (e_k - e_j))"""
    alp = 0.5 * AUTOEV
    t, k = 0.0, i
    for j in range(i + 1, na + 1):
        t += trelax * np.exp(alp * (emo[k - 1] - emo[j - 1]))
        k += 1
    return t


def vary_energies(e_in, e_distr, rng):
    """reference src/boxmuller.f90:46-76 (Box-Muller normal deviate)"""
    dum, dum2 = rng.random(), rng.random()
    z0 = np.sqrt(-2.0 * np.log(dum)) * np.cos(2.0 * np.pi * dum2)
    return e_in + (e_distr * e_in) * z0


def ground_state_sampling(num, mass, xyz, nmax0, tinit=500.0, etemp_gs=298.15, tstep_fs=0.5, mchrg=0, rng=None, device=0, method=gfn2_xtb):
    """The two ground-state runs of src/main.F90:523-567 on the GPU: uniform start velocities for Tinit, equilibration (it = -1,
    nmax0 / 2 steps), sampling (it = 0, nmax0 steps).  Returns dict(records [nmax0, nat, 6], Tav, Epav, Ekav)."""
    rng = np.random.default_rng() if rng is None else rng
    nat = len(num)
    velo = mdinitu(mass, 3.0 * 0.5 * KB * tinit * nat, rng)
    one = np.ones((1, nat))
    eq = Ensemble(num, mass, 1, mchrg=mchrg, tstep_fs=tstep_fs, nmax=max(nmax0 // 2, 1), etemp=etemp_gs, device=device, method=method)
    eq.set_all(np.asarray(xyz)[None], velo[None], one, np.zeros(1), np.zeros(1))
    eq.set_gs_mode(-1, tinit)
    eq.run_md()
    r = eq.result(0)
    eq.close()
    if not r["mdok"]:
        raise RuntimeError("ground-state equilibration failed")
    sm = Ensemble(num, mass, 1, mchrg=mchrg, tstep_fs=tstep_fs, nmax=nmax0, etemp=etemp_gs, device=device, method=method)
    sm.set_all(r["xyz"][None], r["velo"][None], one, np.zeros(1), np.zeros(1))
    sm.set_gs_mode(0)
    sm.run_md()
    r2 = sm.result(0)
    rec = sm.gs_records(0, 0, r2["nstep"])
    sm.close()
    return dict(records=rec, Tav=r2["Tav"], Epav=r2["Epav"], Ekav=r2["Ekav"], mdok=r2["mdok"])


def prepare_runs(num, mass, records, ntraj, rng, eimp0_ev=70.0, eimpw=0.1, ieeatm=0.6, trelax=2000.0, hacc=3.0, fimp=1.0, edistri=1,
                 iee_a=-99.0, iee_b=-99.0, unity=False, etemp_gs=298.15, method=gfn2_xtb):
    """The per-trajectory draw of src/main.F90:684-886 (EI).  records: [ndumpGS, nat, 6] from the sampling run; the orbital data of
    `getspec` come from one spec_calc single point of the neutral molecule at the LAST record, as in the reference (it calls getspec
    with the coordinates left over from reading qcxms.gs, :713).  Returns dict(xyz, velo, velof, eimp, tadd [a.u.], step, mo)."""
    num = np.asarray(num, dtype=np.int32)
    nuc, ndump = len(num), len(records)
    if ndump <= 2 * ntraj:
        raise ValueError("Error: compute longer GS trajectory")
    icalc = np.zeros(ndump + 1, dtype=int)
    k = 0
    while k <= ntraj:
        j = irand(ndump, rng)
        if icalc[j] == 0:
            k += 1
            icalc[j] = 1
    spec = get_xtb_egrad_spec(num, records[-1][:, :3], 0, 1, method, etemp_gs)
    emo, mopop, ihomo = spec["emo"], spec["qmo"], int(spec["ihomo"])
    nb = int(round(spec["focc"].sum())) - ihomo          # beta electrons of the closed-shell neutral = ihomo; getspec returns both
    ehomo = emo[ihomo - 1]
    eimp0 = eimp0_ev * EVTOAU
    exc = (eimp0 - ehomo) * AUTOEV
    ieeel = float(ihomo + nb)
    if not (iee_a > 0 and iee_b > 0):
        iee_a, iee_b = getieeab(ieeel, edistri, exc, nuc, ieeatm)
    _, pmax, _ = getmaxiee(iee_a, iee_b, ieeel, edistri, exc)
    out = dict(xyz=[], velo=[], velof=[], eimp=[], tadd=[], step=[], mo=[], iee_a=iee_a, iee_b=iee_b)
    for i in range(1, ndump + 1):
        if icalc[i] == 0:
            continue
        if len(out["eimp"]) == ntraj:
            break
        while True:
            edum = vary_energies(eimp0, eimpw, rng)
            if edum >= ehomo:
                break
        edum = edum - ehomo
        while True:
            x = rng.random() * edum * AUTOEV
            p = gauss0(iee_a, iee_b, ieeel, x) if edistri == 0 else poiss0(iee_a, iee_b, ieeel, x)
            if p / pmax >= rng.random():
                break
        edum = fimp * x * EVTOAU
        mo1, mo2 = momap(ihomo, emo, edum + ehomo, rng)
        modum = mopop[mo1 - 1].copy()
        if mo2 > 0:
            modum = modum + mopop[mo2 - 1]
        modum[num == 1] *= hacc
        velof = np.ones(nuc) if (nuc > 35 or unity) else modum / modum.max()
        tadd = calctrelax(emo, ihomo, mo1, trelax)
        if mo2 > 0:
            tadd += calctrelax(emo, ihomo, mo2, trelax)
        tadd = max(tadd / (ihomo + nb), trelax / 10.0)
        out["xyz"].append(records[i - 1][:, :3].copy()); out["velo"].append(records[i - 1][:, 3:].copy())
        out["velof"].append(velof); out["eimp"].append(edum); out["tadd"].append(tadd * FSTOAU); out["step"].append(i); out["mo"].append((mo1, mo2))
    for k in ("xyz", "velo", "velof", "eimp", "tadd"):
        out[k] = np.array(out[k])
    out["mass"] = np.asarray(mass, dtype=np.float64)
    return out


def write_directories(root, num, runs):
    """TMPQCXMS/TMP.<n>/{start.xyz, qcxms.start} for every prepared run (src/utility.f90:327-372)"""
    for n in range(len(runs["eimp"])):
        startfiles.write_start(os.path.join(root, "TMPQCXMS", "TMP.%d" % (n + 1)), n + 1, num, runs["xyz"][n], runs["velo"][n], runs["velof"][n],
                               runs["eimp"][n], runs["tadd"][n])
