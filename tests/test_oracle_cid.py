"""CPU tests of the CID restatements (src/diag3x3.f90, src/rotation.f90, src/boxmuller.f90, src/cid.f90)."""
import numpy as np

from qcxms_b200 import ensemble_setup as es
from qcxms_b200.api import AUTOEV, KB, cid_config, load_molecule


def test_eigvec3x3_matches_lapack(oracle):
    rng = np.random.default_rng(5)
    for _ in range(20):
        a = rng.standard_normal((3, 3)); a = a + a.T
        w, q = oracle.eigvec3x3(a.copy())
        assert np.allclose(np.sort(w), np.linalg.eigvalsh(a), atol=1e-12)
        assert np.abs(q.T @ q - np.eye(3)).max() < 1e-12
        assert np.abs(a @ q - q * w).max() < 1e-10
    w, q = oracle.eigvec3x3(np.zeros((3, 3)))
    assert np.all(w == 0) and np.array_equal(q, np.eye(3))


def test_euler_rotation_is_a_rigid_rotation(oracle):
    rng = np.random.default_rng(6)
    xyz = rng.standard_normal((8, 3)); velo = rng.standard_normal((8, 3))
    x2, v2 = oracle.euler_rotation(xyz, velo, 0.3, 0.8, 0.55)
    assert np.allclose(np.linalg.norm(x2, axis=1), np.linalg.norm(xyz, axis=1))
    assert np.allclose(x2 @ x2.T, xyz @ xyz.T) and np.allclose(x2 @ v2.T, xyz @ velo.T)
    x3, _ = oracle.euler_rotation(xyz, velo, 0.0, 0.0, 0.0)
    assert np.allclose(x3, xyz)
    # R = R_alpha(x) R_beta(y) R_gamma(z): a pure alpha rotation leaves x untouched
    x4, _ = oracle.euler_rotation(xyz, velo, 0.25, 0.0, 0.0)
    assert np.allclose(x4[:, 0], xyz[:, 0]) and np.allclose(x4[:, 1], -xyz[:, 2]) and np.allclose(x4[:, 2], xyz[:, 1])


def test_rotation_velo_energy(oracle):
    """E_rot = sum_k 1/2 I_k w_k^2 with w_k = sqrt(kB T / I_k) -> 3/2 kB T (reference src/rotation.f90:150-175)."""
    num, xyz, _ = load_molecule("chloroethanol")
    mass = es.masses_au(num)
    com = (mass[:, None] * xyz).sum(0) / mass.sum()
    rng = np.random.default_rng(7)
    velo = 1e-4 * rng.standard_normal(xyz.shape)
    vrot, erot = oracle.rotation_velo(xyz - com, mass, velo)
    _, T = oracle.ekinet(velo, mass)
    assert abs(erot - 1.5 * KB * T) < 1e-15
    assert np.isfinite(vrot).all() and np.abs(vrot).max() > 0


def test_vary_energies_box_muller(oracle):
    e = oracle.vary_energies(40.0, 0.1, 0.7, 0.25)    # dum > 0.5: z0 = sqrt(-2 ln dum) cos(2 pi dum2) = 0 at dum2 = 1/4
    assert abs(e - 40.0) < 1e-12
    e = oracle.vary_energies(40.0, 0.1, 0.3, 0.25)    # dum <= 0.5: z1 = sqrt(-2 ln dum) sin(pi/2)
    assert abs(e - (40.0 + 4.0 * np.sqrt(-2 * np.log(0.3)))) < 1e-12
    rng = np.random.default_rng(8)
    s = np.array([oracle.vary_energies(40.0, 0.1, *rng.random(2)) for _ in range(4000)])
    assert abs(s.mean() - 40.0) < 0.3 and abs(s.std() - 4.0) < 0.3


def test_cid_first_collision_setup_and_short_loop(oracle):
    num, xyz, _ = load_molecule("chloroethanol")
    mass = es.masses_au(num)
    ic = es.synthetic_initial_conditions(num, xyz, 1, first_id=2)
    cfg = cid_config(mchrg=1, gas="ar", elab=40.0, ntot=6, eexact=True)
    rnd = np.array([0.11, 0.63, 0.42, 0.7, 0.2, 0.5, 0.25, 0.8, 0.1])
    out = oracle.cid(cfg, num, mass, 1, ic["xyz"][0], ic["velo"][0], rnd)
    assert out["status"] == 1 and out["stopcid"] == 0 and out["nstep"] == 6 and out["nfrag"] == 1 and out["collided"] == 0
    assert np.array_equal(out["list"], np.ones(len(num), dtype=np.int32))
    # the ion flies along direc with the laboratory-frame speed sqrt(2 E_lab / M)
    v_expect = np.sqrt(2 * 40.0 / AUTOEV / mass.sum())
    vcm = (mass[:, None] * out["velo"]).sum(0) / mass.sum()
    # (on top of the thermal centre-of-mass drift of the initial conditions, a few per cent)
    assert abs(np.linalg.norm(vcm) - v_expect) / v_expect < 5e-2
    assert np.allclose(vcm / np.linalg.norm(vcm), out["direc"] / np.linalg.norm(out["direc"]), atol=5e-2)
    assert abs(np.linalg.norm(out["direc"]) - 1.0) < 0.05        # sequential normalisation of the reference (sic)
    assert abs(out["velo_cm"] * (1.0 / 2.18769126364e+06) - np.linalg.norm(vcm)) / v_expect < 1e-3
    assert out["scc_iter_total"] > 7 * 5


def test_cid_later_collision_keeps_velocities(oracle):
    num, xyz, _ = load_molecule("chloroethanol")
    mass = es.masses_au(num)
    ic = es.synthetic_initial_conditions(num, xyz, 1, first_id=3)
    cfg = cid_config(mchrg=1, gas="ar", elab=40.0, ntot=3, eexact=True)
    rnd = np.array([0.3, 0.3, 0.3, 0.6, 0.6, 0.2, 0.9, 0.4, 0.6])
    direc = np.array([0.0, 0.0, 1.0])
    out = oracle.cid(cfg, num, mass, 2, ic["xyz"][0], ic["velo"][0], rnd, velo_cm=9000.0, direc=direc, collided=1)
    assert out["status"] == 1 and out["nstep"] == 3 and out["collided"] == 1
    # no lab-frame boost is added for icoll > 1: the centre-of-mass velocity is the one that was handed in
    vcm0 = (mass[:, None] * ic["velo"][0]).sum(0) / mass.sum()
    vcm1 = (mass[:, None] * out["velo"]).sum(0) / mass.sum()
    assert np.abs(vcm1 - vcm0).max() < 1e-6   # the gas atom 17 bohr away pulls a little
    assert np.array_equal(out["direc"], direc)


def test_cid_n2_has_two_gas_atoms(oracle):
    """N2 (gas_z 7) adds two atoms: one more than Ar in the single points (more SCC work is not observable, the step count
    and a clean exit are), and the ion still leaves with the laboratory-frame speed."""
    num, xyz, _ = load_molecule("chloroethanol")
    mass = es.masses_au(num)
    ic = es.synthetic_initial_conditions(num, xyz, 1, first_id=4)
    rnd = np.array([0.21, 0.33, 0.72, 0.6, 0.3, 0.4, 0.35, 0.2, 0.7])
    cfg = cid_config(mchrg=1, gas="n2", elab=30.0, ntot=5, eexact=True)
    assert cfg.gas_z == 7 and abs(cfg.gas_mass - 14.007 * 1822.888486) / cfg.gas_mass < 1e-6
    out = oracle.cid(cfg, num, mass, 1, ic["xyz"][0], ic["velo"][0], rnd)
    assert out["status"] == 1 and out["nstep"] == 5 and out["nfrag"] == 1
    v_expect = np.sqrt(2 * 30.0 / AUTOEV / mass.sum())
    vcm = (mass[:, None] * out["velo"]).sum(0) / mass.sum()
    assert abs(np.linalg.norm(vcm) - v_expect) / v_expect < 5e-2
