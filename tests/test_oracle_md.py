"""CPU tests of the MD-side restatements (leapfrog, impactscale, fragments, checkqc, md state machine)."""
import json
import os

import numpy as np

from qcxms_b200 import ensemble_setup as es
from qcxms_b200.api import FSTOAU, KB, load_molecule

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_leapfrog_and_ekinet(oracle):
    rng = np.random.default_rng(0)
    n = 7
    mass = rng.uniform(1800, 40000, n); xyz = rng.standard_normal((n, 3)); v = 1e-3 * rng.standard_normal((n, 3)); g = 1e-2 * rng.standard_normal((n, 3))
    x2, v2, ke = oracle.leapfrog(g, mass, 20.0, xyz, v)
    vn = v - 20.0 * g / mass[:, None]
    assert np.allclose(v2, vn, rtol=0, atol=1e-18) and np.allclose(x2, xyz + 20.0 * vn, rtol=0, atol=1e-15)
    assert abs(ke - (0.5 * mass[:, None] * (0.5 * (v + vn)) ** 2).sum()) < 1e-14
    e, t = oracle.ekinet(v, mass)
    assert abs(e - 0.5 * (mass[:, None] * v ** 2).sum()) < 1e-15 and abs(t - e / (1.5 * n * KB)) < 1e-9


def test_impactscale_grid_and_off_by_one(oracle):
    """reference src/impact.f90:31-48: scal runs over multiples of the single-precision literal 0.0002 and is
    incremented once more after the first passing test."""
    rng = np.random.default_rng(1)
    n = 5
    mass = rng.uniform(1800, 30000, n); v = 3e-4 * rng.standard_normal((n, 3)); velof = np.ones(n)
    e0 = oracle.ekinet(v, mass)[0]
    step = float(np.float32(0.0002))
    # target reachable without scaling: the loop still applies one grid step
    v1, err = oracle.impactscale(v, mass, velof, eimp=0.0, ff=1.0, e0=e0)
    assert err == 0 and np.allclose(v1, v * (1 + step), rtol=0, atol=1e-20)
    # a real heating target: first k with Esoll - E(k*step) <= 0.001f, applied factor (k+1)*step
    target = e0 * 1.5 + 0.01
    v2, err = oracle.impactscale(v, mass, velof, eimp=target - e0, ff=1.0, e0=e0)
    k = 0
    while target - e0 * (1 + k * step) ** 2 > float(np.float32(0.001)):
        k += 1
    assert err == 0 and np.allclose(v2, v * (1 + (k + 1) * step), rtol=1e-12, atol=0)
    # velof = 0 can never heat: the reference stops with an error after 20000 trials
    _, err = oracle.impactscale(v, mass, np.zeros(n), eimp=1.0, ff=1.0, e0=e0)
    assert err == 1


def test_fragment_structure_hand_built(oracle):
    # H2 ... H2 far apart ... lone H : labels follow the lowest atom index of each fragment (src/fragments.f90:131-178)
    num = np.array([1, 1, 1, 1, 1])
    xyz = np.array([[0, 0, 0], [30, 0, 0], [1.4, 0, 0], [31.4, 0, 0], [0, 50, 0]], dtype=float)
    assert oracle.fragment_structure(num, xyz).tolist() == [1, 2, 1, 2, 3]
    # threshold r < rcut * 0.5 * (Rad_i + Rad_j), Rad_H = 0.32 / 0.52917726
    r0 = 3.0 * 0.32 / 0.52917726
    assert oracle.fragment_structure(num[:2], np.array([[0, 0, 0], [0, 0, np.nextafter(r0, 0)]])).tolist() == [1, 1]
    assert oracle.fragment_structure(num[:2], np.array([[0, 0, 0], [0, 0, r0 * (1 + 1e-12)]])).tolist() == [1, 2]
    # at1 = at2 = 0 deletes the assignment
    assert oracle.fragment_structure(num, xyz, at1=0, at2=0).tolist() == [1] * 5
    # intact molecule
    n2, x2, _ = load_molecule("caffeine")
    assert set(oracle.fragment_structure(n2, x2).tolist()) == {1}


def test_fragmass(oracle):
    num = np.array([6, 1, 1, 17, 1]); lst = np.array([1, 1, 2, 2, 3])
    mass = es.masses_au(num)
    nfrag, fragx, fragat = oracle.fragmass(num, lst, mass)
    assert nfrag == 3
    assert abs(fragx[0] - (12.0107359 + 1.00794075)) < 1e-6 and abs(fragx[1] - (35.45293758 + 1.00794075)) < 1e-6
    assert fragat[0][6 - 1] == 1 and fragat[0][1 - 1] == 1 and fragat[1][17 - 1] == 1 and fragat[2][0] == 1
    # isotope labelling: j = 100 + imass
    nfrag, _, fragat = oracle.fragmass(num, lst, mass, imass=np.array([13, -1, -1, -1, 2]))
    assert fragat[0][113 - 1] == 1 and fragat[2][102 - 1] == 1


def test_checkqc_quirks(oracle):
    q = np.array([0.5, 0.5])
    g = np.array([[0.0, 0.0, 5.0], [0.0, 0.0, -5.0]])
    # gnorm counts g_y twice and never g_z (src/iniqm.f90:727): a pure-z gradient looks like zero -> rejected
    ok, e = oracle.checkqc(-10.0, g, q, 1)
    assert not ok and e == 0.0
    g2 = np.array([[0.0, 20.0 / np.sqrt(2) * 1.001, 0.0], [0.0, 0.0, 0.0]])
    assert oracle.checkqc(-10.0, g2, q, 1)[0] is False          # sqrt(2 gy^2) > 20
    assert oracle.checkqc(-10.0, g2 * 0.99, q, 1)[0] is True
    assert oracle.checkqc(1e-9, g2 * 0.5, q, 1) == (False, 1e-9)   # |E| < 1e-8: rejected, E untouched
    assert oracle.checkqc(-10.0, g2 * 0.5, np.zeros(2), 1)[0] is False and oracle.checkqc(-10.0, g2 * 0.5, np.zeros(2), 0)[0] is True


def test_setetemp_getspin(oracle):
    assert oracle.setetemp(1, 0.5) == 5000.0 and oracle.setetemp(1, 0.5, ax=0.25) == 10000.0
    assert oracle.setetemp(1, 0.5, ieetemp=1000.0) == 5500.0 and oracle.setetemp(2, 0.5, ieetemp=1000.0) == 5000.0
    assert oracle.setetemp(1, float("nan"), ieetemp=1000.0) == 5000.0      # nadd = 0 in secondary runs
    assert oracle.getspin([6, 1, 1, 1, 1], 0) == 1 and oracle.getspin([6, 1, 1, 1, 1], 1) == 2 and oracle.getspin([1], 1) == -1


def test_md_golden_and_energy_conservation(oracle):
    num, xyz, _ = load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 2)
    gold = json.load(open(os.path.join(GOLD, "md_golden.json")))
    for c in gold["cases"]:
        k = c["traj"]
        r = oracle.md(num, ic["mass"], ic["xyz"][k], ic["velo"][k], ic["velof"][k], ic["eimp"][k], ic["tadd"][k], mchrg=1, nmax=10)
        assert r["nstep"] == c["nstep"] == 10 and r["mdok"] == 1 and r["fragstate"] == 1 and r["scc_iter_total"] == c["scc_iter_total"]
        assert np.abs(r["xyz"] - np.array(c["xyz"])).max() < 1e-10 and np.abs(r["velo"] - np.array(c["velo"])).max() < 1e-12
        assert r["list"].tolist() == c["list"]
    # NVE (no IEE: velof = 0 is never used because tadd = 0 -> nadd = 0): total energy is conserved over 40 steps
    r0 = oracle.md(num, ic["mass"], ic["xyz"][0], ic["velo"][0], ic["velof"][0], 0.0, 0.0, mchrg=1, nmax=1)
    r = oracle.md(num, ic["mass"], ic["xyz"][0], ic["velo"][0], ic["velof"][0], 0.0, 0.0, mchrg=1, nmax=40)
    assert abs((r["Epot"] + r["Ekin"]) - (r0["Epot"] + r0["Ekin"])) < 1e-3


def test_user_etemp_serves_only_the_first_single_point(oracle):
    """src/md.f90:167-172 vs 443-445: etempin >= 0 sets the electronic temperature of the initial egrad; inside the loop setetemp
    is called on every step regardless, so from step 1 on the run is the default-temperature run."""
    from qcxms_b200 import ensemble_setup as es
    from qcxms_b200.api import load_molecule
    num, xyz, _ = load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 1, first_id=77)
    a = (ic["mass"], ic["xyz"][0], ic["velo"][0], ic["velof"][0], ic["eimp"][0], ic["tadd"][0])
    user = oracle.md(num, *a, mchrg=1, nmax=3, etemp=300.0)
    dflt = oracle.md(num, *a, mchrg=1, nmax=3, etemp=-1.0)
    # the first gradient (etemp 300 K vs 5000 K + IEE term) differs, so the trajectories differ slightly ...
    assert 1e-9 < np.abs(user["xyz"] - dflt["xyz"]).max() < 1e-2
    # ... but the last single point was evaluated at the setetemp temperature, not at 300 K
    e5000 = oracle.egrad(num, user["xyz"], charge=1, multiplicity=2, etemp=5000.0)["energy"]
    e300 = oracle.egrad(num, user["xyz"], charge=1, multiplicity=2, etemp=300.0)["energy"]
    assert abs(user["Epot"] - e5000) < 1e-9 < abs(user["Epot"] - e300)
