"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances are the ones BASELINE.json's north_star states: energy 1e-8 Eh, gradient 1e-6 Eh/bohr,
Mulliken charges 1e-6; fragment ids bit-exact.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

E_TOL, G_TOL, Q_TOL = 1e-8, 1e-6, 1e-6


def _compare(qx, oracle, num, xyz, charge, mult, etemp):
    q, e, g, stat = qx.get_xtb_egrad(num, xyz, charge, mult, qx.gfn2_xtb, etemp)
    ref = oracle.egrad(num, xyz, charge=charge, multiplicity=mult, etemp=etemp, detail=True)
    assert stat == ref["stat"] == 0
    assert abs(e - ref["energy"]) < E_TOL, (e, ref["energy"])
    assert np.abs(g - ref["gradient"]).max() < G_TOL
    assert np.abs(q - ref["qat"]).max() < Q_TOL
    return ref


@pytest.mark.parametrize("name,charge,mult,etemp", [
    ("chloroethanol", 0, 1, 300.0), ("chloroethanol", 1, 2, 5000.0), ("monoethanolamine", 1, 2, 5000.0),
    ("thf_h", 1, 1, 5000.0), ("dichlorobenzamide_h", 1, 1, 5000.0), ("caffeine", 1, 2, 5000.0), ("caffeine", 0, 1, 300.0)])
def test_egrad_matches_oracle(qx, oracle, name, charge, mult, etemp):
    num, xyz, _ = qx.load_molecule(name)
    _compare(qx, oracle, num, xyz, charge, mult, etemp)


def test_egrad_batch_distorted_geometries(qx, oracle):
    num, xyz, _ = qx.load_molecule("caffeine")
    rng = np.random.default_rng(7)
    geoms = xyz[None] + 0.08 * rng.standard_normal((12,) + xyz.shape)
    out = qx.egrad_batch(num, geoms, 1, 2, qx.gfn2_xtb, 5000.0)
    for k in range(len(geoms)):
        ref = oracle.egrad(num, geoms[k], charge=1, multiplicity=2, etemp=5000.0, detail=True)
        assert out["stat"][k] == 0
        assert out["niter"][k] == ref["niter"]
        assert abs(out["energy"][k] - ref["energy"]) < E_TOL
        assert np.abs(out["gradient"][k] - ref["gradient"]).max() < G_TOL
        assert np.abs(out["qat"][k] - ref["qat"]).max() < Q_TOL


def test_unknown_method_sets_stat_5(qx):
    num, xyz, _ = qx.load_molecule("chloroethanol")
    _, _, _, stat = qx.get_xtb_egrad(num, xyz, 0, 1, 99, 300.0)
    assert stat == 5
    _, _, _, stat = qx.get_xtb_egrad(num, xyz, 0, 1, qx.ipea1_xtb, 300.0)   # IPEA1: GFN1 model, element table not reconstructed
    assert stat == 5


@pytest.mark.parametrize("name,charge,mult,etemp", [
    ("monoethanolamine", 0, 1, 300.0), ("monoethanolamine", 1, 2, 5000.0), ("caffeine", 1, 2, 5000.0), ("caffeine", 0, 1, 300.0),
    ("chloroethanol", 1, 2, 5000.0), ("xb", 0, 1, 300.0)])
def test_gfn1_egrad_matches_oracle(qx, oracle, name, charge, mult, etemp):
    """GFN1-xTB (method id 1; BASELINE config 3): exponential CN, two s shells on H, D3(BJ), halogen bond, atomic third order."""
    if name == "xb":
        from test_oracle_gfn1 import XB_NUM as num, XB_XYZ as xyz
    else:
        num, xyz, _ = qx.load_molecule(name)
        xyz = xyz + 0.04 * np.random.default_rng(11).standard_normal(xyz.shape)
    q, e, g, stat = qx.get_xtb_egrad(num, xyz, charge, mult, qx.gfn1_xtb, etemp)
    ref = oracle.egrad(num, xyz, charge=charge, multiplicity=mult, method=1, etemp=etemp, detail=True)
    assert stat == ref["stat"] == 0
    assert abs(e - ref["energy"]) < E_TOL, (e, ref["energy"])
    assert np.abs(g - ref["gradient"]).max() < G_TOL
    assert np.abs(q - ref["qat"]).max() < Q_TOL


def test_gfn1_batch_niter_and_md(qx, oracle):
    from qcxms_b200 import ensemble_setup as es
    num, xyz, _ = qx.load_molecule("monoethanolamine")
    geoms = xyz[None] + 0.08 * np.random.default_rng(7).standard_normal((8,) + xyz.shape)
    out = qx.egrad_batch(num, geoms, 1, 2, qx.gfn1_xtb, 5000.0)
    for k in range(len(geoms)):
        ref = oracle.egrad(num, geoms[k], charge=1, multiplicity=2, method=1, etemp=5000.0, detail=True)
        assert out["stat"][k] == 0 and out["niter"][k] == ref["niter"]
        assert abs(out["energy"][k] - ref["energy"]) < E_TOL and np.abs(out["gradient"][k] - ref["gradient"]).max() < G_TOL
    # md() with the GFN1 calculator: a short EI run against the oracle's md
    ic = es.synthetic_initial_conditions(num, xyz, 3)
    ens = qx.Ensemble(num, ic["mass"], 3, mchrg=1, nmax=12, exit_rules=True, method=qx.gfn1_xtb)
    ens.set_all(ic["xyz"], ic["velo"], ic["velof"], ic["eimp"], ic["tadd"])
    assert ens.run_md() == 3 * 12
    for k in range(3):
        got = ens.result(k)
        ref = oracle.md(num, ic["mass"], ic["xyz"][k], ic["velo"][k], ic["velof"][k], ic["eimp"][k], ic["tadd"][k], mchrg=1, nmax=12, method=1)
        assert got["nstep"] == ref["nstep"] == 12 and got["scc_iter_total"] == ref["scc_iter_total"]
        assert np.abs(got["xyz"] - ref["xyz"]).max() < 1e-7 and abs(got["Epot"] - ref["Epot"]) < 1e-7
        assert np.array_equal(got["list"], ref["list"])
    ens.close()


def test_large_basis_fallback_alkane_c32(qx, oracle):
    """98 atoms / 194 AOs: the SCC matrices no longer fit shared memory (functional global-memory fallback path)."""
    num, xyz, _ = qx.load_molecule("alkane_c32")
    rng = np.random.default_rng(2)
    x = xyz + 0.03 * rng.standard_normal(xyz.shape)
    _compare(qx, oracle, num, x, 1, 2, 5000.0)


def test_config5_drug_like_peptide(qx, oracle):
    """BASELINE config 5: 98 atoms with N, O, S and Cl (d functions in the large-basis path), 152 shells / 266 AOs -- above the 256 the
    blocked Jacobi used to be capped at.  Relaxed neutral geometry and a distorted cation."""
    num, xyz, _ = qx.load_molecule("peptide_cl")
    assert len(num) == 98 and set(num.tolist()) == {1, 6, 7, 8, 16, 17}
    ref = _compare(qx, oracle, num, xyz, 0, 1, 300.0)
    assert ref["nao"] == 266
    rng = np.random.default_rng(5)
    _compare(qx, oracle, num, xyz + 0.03 * rng.standard_normal(xyz.shape), 1, 2, 5000.0)


@pytest.mark.parametrize("name", ["alkane_c14", "alkane_c17"])
def test_medium_basis_multi_pass_jacobi(qx, oracle, name):
    """Bases between the one-pass Jacobi limit (72 AOs) and the shared-memory limit (~110): matrices in shared memory,
    generic DMMA GEMMs, multi-pass Jacobi (jacobi_rows_lp8m)."""
    num, xyz, _ = qx.load_molecule(name)
    rng = np.random.default_rng(2)
    x = xyz + 0.03 * rng.standard_normal(xyz.shape)
    _compare(qx, oracle, num, x, 1, 2, 5000.0)


@pytest.mark.parametrize("name,charge,etemp", [("chloroethanol", 0, 300.0), ("chloroethanol", 1, 5000.0), ("thf_h", 1, 300.0)])
def test_spec_calc_output_matches_oracle(qx, oracle, tmp_path, name, charge, etemp):
    """get_xtb_egrad(..., spec_calc = .true.) (reference src/tblite.f90:152-164, write_qmo src/mo_energ.f90:7-80): orbital energies,
    occupations, HOMO index and the normalised atomic populations of every orbital; energy / gradient / charges unchanged."""
    from qcxms_b200.fragments import getspin
    num, xyz, _ = qx.load_molecule(name)
    mult = getspin(num, charge)
    ref = oracle.egrad(num, xyz, charge=charge, multiplicity=mult, etemp=etemp, detail=True)
    qref = oracle.qmo(ref["ao2at"], ref["coeff"], ref["overlap"], len(num))
    got = qx.get_xtb_egrad_spec(num, xyz, charge, mult, qx.gfn2_xtb, etemp, write_files=str(tmp_path))
    pq, pe, pg, pstat = qx.get_xtb_egrad(num, xyz, charge, mult, qx.gfn2_xtb, etemp)
    plain = dict(qat=pq, energy=pe, gradient=pg)
    assert got["stat"] == 0 and pstat == 0 and got["nao"] == ref["nao"] and got["ihomo"] == ref["ihomo"]
    assert got["energy"] == plain["energy"] and np.array_equal(got["gradient"], plain["gradient"]) and np.array_equal(got["qat"], plain["qat"])
    assert np.abs(got["emo"] - ref["emo"]).max() < 1e-6          # Eh; both stop at the SCC thresholds
    assert np.abs(got["focc"] - ref["focc"]).max() < 1e-5
    assert np.all(np.diff(got["emo"]) >= 0) and abs(got["focc"].sum() - ref["focc"].sum()) < 1e-9
    assert np.abs(got["qmo"].sum(1) - 1.0).max() < 1e-12
    gaps = np.diff(ref["emo"])
    ok = np.concatenate(([True], gaps > 1e-3)) & np.concatenate((gaps > 1e-3, [True]))     # populations of degenerate levels are not unique
    assert ok.sum() >= ref["nao"] - 4
    assert np.abs(got["qmo"][ok] - qref[ok]).max() < 1e-4
    # the two files getspec reads (src/mo_spec.f90): header and record count
    tmp = (tmp_path / "tmp.mspec").read_text().split()
    assert int(tmp[0]) == ref["nao"] and int(tmp[1]) == ref["ihomo"] and len(tmp) == 2 + ref["nao"] * (2 + len(num))
    assert abs(float(tmp[2]) - got["emo"][0] * 27.21138505) < 1e-9
    assert (tmp_path / "qcxms.Mspec.tbxtb").read_text().split()[0] == str(ref["nao"])
