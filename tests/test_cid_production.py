"""CID production run (reference main.F90:1490-2163): mean-free-path md() restatement and the collision loop on the CPU oracle back end."""
import numpy as np

from qcxms_b200 import ensemble_setup as es
from qcxms_b200 import production as prod
from qcxms_b200.api import KB, load_molecule

MSTOAU = 1.0 / 2.18769126364e+06


def test_mfp_md_removes_the_centre_of_mass_motion(oracle):
    """md() with method 3, icoll >= 1 (src/md.f90:246-255, 466-493): temperature and kinetic energy are those of the internal
    motion, so a Galilean boost of the ion changes new_velo and nothing else."""
    num, xyz, _ = load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 1, first_id=7)
    mass, x0, v0 = ic["mass"], ic["xyz"][0], ic["velo"][0]
    vcm = (mass[:, None] * v0).sum(0) / mass.sum()
    v_int = v0 - vcm                                             # no drift at all
    boost = np.array([0.0, 0.0, 9000.0 * MSTOAU])                # 9 km/s along z
    a = oracle.md_mfp(num, mass, x0, v_int, 1, 0.0, nmax=12)
    b = oracle.md_mfp(num, mass, x0, v_int + boost, 1, 9000.0, nmax=12)
    assert a["nstep"] == b["nstep"] == 12 and a["mdok"] == b["mdok"] == 1 and a["fragstate"] == 1 and a["nfrag"] == 1
    assert abs(a["new_velo"]) < 1e-3 and abs(b["new_velo"] - 9000.0) < 1e-3        # m/s
    assert abs(a["Tav"] - b["Tav"]) < 1e-3 * a["Tav"] and abs(a["Ekin"] - b["Ekin"]) < 1e-6
    assert abs(a["Epav"] - b["Epav"]) < 1e-9
    assert np.abs((b["xyz"] - a["xyz"]) - 12 * 0.5 * 41.3413733365614 * boost).max() < 1e-7
    # aTlast is the mean of new_temp, the internal temperature
    _, T = oracle.ekinet(v_int, mass)
    assert abs(a["aTlast"] - T) < 0.25 * T
    # no IEE heating in this mode: the same call through the EI branch with eimp = 0 differs only by the averaging bookkeeping
    c = oracle.md(num, mass, x0, v_int, np.ones(len(num)), 0.0, 0.0, nmax=12, etemp=-1.0, isec=2)
    assert np.abs(c["xyz"] - a["xyz"]).max() < 1e-12 and abs(c["Epot"] - a["Epot"]) < 1e-12


def test_mfp_md_counts_steps_after_a_fragmentation(oracle):
    """A fragmenting ion: the end of the run moves to nstep + add_steps (0 for nuc <= 10: src/md.f90:233, 507, 672) and axyz is the
    averaged structure of the counted steps (src/md.f90:526-621, 694-699)."""
    num, xyz, _ = load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 1, first_id=8)
    mass, x0 = ic["mass"], ic["xyz"][0]
    v = ic["velo"][0] * 1.0
    icl = int(np.nonzero(num == 17)[0][0])
    ic_c = int(np.argmin(np.where(num == 6, np.linalg.norm(x0 - x0[icl], axis=1), 1e9)))
    d = (x0[icl] - x0[ic_c]) / np.linalg.norm(x0[icl] - x0[ic_c])
    v[icl] += 2.5e-3 * d                                         # shoot the chlorine away
    out = oracle.md_mfp(num, mass, x0, v, 1, 0.0, nmax=400)
    assert out["nfrag"] == 2 and out["mdok"] == 1 and out["fragstate"] == 1
    assert out["nstep"] < 400                                    # ended at the step of the fragmentation (add_steps = 0 for 9 atoms)
    # with the run ending on the first counted step, store_avxyz is that single structure
    assert np.abs(out["axyz"] - out["xyz"]).max() < 1e-12
    assert sorted(np.bincount(out["list"])[1:].tolist()) == [1, 8]


def test_collision_numbers():
    num, xyz, _ = load_molecule("dichlorobenzamide_h")
    r_mol, cross, mfpath, ncoll = prod.collision_setup(num, xyz, "ar", lchamb=0.25)
    assert 3e-10 < r_mol < 5e-10 and abs(cross - np.pi * (r_mol + 3.55266638 * 0.52917726e-10) ** 2) < 1e-30
    assert abs(mfpath - 1.38064852e-23 * 300.0 / (cross * 0.132)) < 1e-12 and abs(ncoll - 0.25 / mfpath) < 1e-12
    assert 5 < ncoll < 12
    # Box-Muller: z0 = 0 at dum2 = 1/4 -> nint(calc_collisions); never negative
    assert prod.vary_collisions(8.5, 0.3, 0.25) == 8 + (1 if 8.5 - 8 >= 0.5 else 0)
    assert prod.vary_collisions(0.4, 0.9, 0.5) == 0
    rng = np.random.default_rng(3)
    n = np.array([prod.vary_collisions(10.0, *rng.random(2)) for _ in range(3000)])
    assert abs(n.mean() - 10.0) < 0.15 and abs(n.std() - 1.2) < 0.15


def test_cid_run_on_oracle_backend(oracle):
    """Two collisions with the mean-free-path MDs between them (maxcoll run type), short runs."""
    num, xyz, _ = load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 2, first_id=300)
    out = prod.run_cid(num, ic["mass"], ic["xyz"], ic["velo"], mchrg=1, gas="ar", elab=40.0, run_type="maxcoll", max_coll=2, minmass=20,
                       first_itrj=5, seed=11, cid_ntot=8, mfp_nmax=6, cid_batch=oracle.cid_batch, mfp_batch=oracle.mfp_batch,
                       energies=oracle.energies)
    assert len(out["per_traj"]) == 2
    for t in out["per_traj"]:
        kinds = [(e["kind"], e["icoll"]) for e in t["events"]]
        assert kinds == [("cid", 1), ("mfp", 1), ("cid", 2), ("mfp", 2)]
        assert all(e["nstep"] == (8 if e["kind"] == "cid" else 6) for e in t["events"])
        v = [e["velo_cm"] if e["kind"] == "cid" else e["new_velo"] for e in t["events"]]
        v_lab = np.sqrt(2 * 40.0 / 27.21138505 / ic["mass"].sum()) / MSTOAU
        assert all(abs(x - v_lab) / v_lab < 0.1 for x in v)      # the ion keeps flying with the laboratory-frame speed
        # nothing fragmented: one record, the intact ion with the full charge, written when the last collision was done
        assert len(t["records"]) == 1 and abs(float(t["records"][0][:10]) - 1.0) < 1e-6
        assert int(t["records"][0][13:18]) == t["itrj"]


def test_cid_run_is_independent_of_batching(oracle):
    num, xyz, _ = load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 2, first_id=310)
    kw = dict(mchrg=1, gas="ar", elab=40.0, run_type="collno", collno=(1, 1, 1), minmass=20, seed=4, cid_ntot=5, mfp_nmax=4,
              cid_batch=oracle.cid_batch, mfp_batch=oracle.mfp_batch, energies=oracle.energies)
    both = prod.run_cid(num, ic["mass"], ic["xyz"], ic["velo"], first_itrj=1, **kw)
    second = prod.run_cid(num, ic["mass"], ic["xyz"][1:], ic["velo"][1:], first_itrj=2, **kw)
    assert both["per_traj"][1]["records"] == second["per_traj"][0]["records"]
    assert both["per_traj"][1]["events"] == second["per_traj"][0]["events"]


def test_n2_stop_rule_uses_the_single_atom_gas_mass(oracle):
    """E_COM = beta E_kin with beta = mIatom / (mIatom + M), mIatom = 14.007 amu also for N2 (main.F90:1993-2000, input.f90
    'IATOM N2'): at E_lab = 5 eV a chloroethanol ion has E_COM = 0.74 eV <= 0.85 eV and the run ends after the first collision;
    with the molecular mass (28 amu) it would be 1.29 eV and a second collision would follow."""
    num, xyz, _ = load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 1, first_id=320)
    out = prod.run_cid(num, ic["mass"], ic["xyz"], ic["velo"], mchrg=1, gas="n2", elab=5.0, run_type="maxcoll", max_coll=3, minmass=20,
                       first_itrj=1, seed=2, cid_ntot=6, mfp_nmax=4, cid_batch=oracle.cid_batch, mfp_batch=oracle.mfp_batch,
                       energies=oracle.energies)
    t = out["per_traj"][0]
    assert [(e["kind"], e["icoll"]) for e in t["events"]] == [("cid", 1), ("mfp", 1)]
    e_kin_ev = 0.5 * ic["mass"].sum() * (t["events"][-1]["new_velo"] * MSTOAU) ** 2 * 27.21138505
    m1 = 14.007 * 1822.888486
    assert m1 / (m1 + ic["mass"].sum()) * e_kin_ev <= 0.85 < 2 * m1 / (2 * m1 + ic["mass"].sum()) * e_kin_ev
    assert len(t["records"]) == 1


def test_esi_preheating_before_the_collisions(oracle):
    """keyword `esi <eV>` (main.F90:1243-1362): ions colder than the requested internal energy are heated by the thermostatted MD
    (md() with method 3, icoll 0, starting_md) before the first collision; ions that are already hotter are left alone."""
    num, xyz, _ = load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 2, first_id=330, temperature=300.0)
    nuc = len(num)
    e_int = [0.5 * (ic["mass"][:, None] * ic["velo"][k] ** 2).sum() * 27.21138505 for k in range(2)]
    kw = dict(mchrg=1, gas="ar", elab=40.0, run_type="maxcoll", max_coll=1, minmass=20, seed=5, cid_ntot=5, mfp_nmax=4, cid_batch=oracle.cid_batch,
              mfp_batch=oracle.mfp_batch, esi_batch=oracle.esi_batch, energies=oracle.energies)
    out = prod.run_cid(num, ic["mass"], ic["xyz"], ic["velo"], esi_ev=2.0 * max(e_int), esi_nmax=12, **kw)
    for t in out["per_traj"]:
        kinds = [(e["kind"], e["icoll"]) for e in t["events"]]
        assert kinds == [("esi", 0), ("cid", 1), ("mfp", 1)] and t["events"][0]["nstep"] == 12 and t["events"][0]["md_ok"]
        tscale = t["events"][0]["tscale"]
        assert abs(tscale - (2.0 * max(e_int) * 2.0 / 3.0) / (nuc * 3.166808578545117e-06 * 27.21138505)) < 1e-9
    # length of the heating run when it is not overridden: nint(2.5 (T_target - T)), main.F90:1335
    cold = prod.run_cid(num, ic["mass"], ic["xyz"][:1], ic["velo"][:1], esi_ev=1.02 * e_int[0], **kw)
    t0 = e_int[0] / 27.21138505 / (1.5 * nuc * 3.166808578545117e-06)
    assert cold["per_traj"][0]["events"][0]["nstep"] == int(np.floor(0.02 * t0 * 2.5 + 0.5))
    # "! No Scaling !": the ion already holds more than the requested energy
    none = prod.run_cid(num, ic["mass"], ic["xyz"][:1], ic["velo"][:1], esi_ev=0.5 * e_int[0], **kw)
    assert [e["kind"] for e in none["per_traj"][0]["events"]] == ["cid", "mfp"]
