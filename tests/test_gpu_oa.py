"""Wide-CTA kernels (576 threads, one CTA per SM) with the GEMM-based eigenpair refinement (csrc/qx_oa.cuh): the eigen-solver
changes, the SCC protocol and its results must not.  Ensembles with at most one trajectory per SM select these kernels on their own;
QCXMS_B200_CTA=576 forces them for larger ones, QCXMS_B200_OA=0 keeps the one-sided Jacobi inside them."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _md(qx, num, ic, n, nsteps):
    ens = qx.Ensemble(num, ic["mass"], n, mchrg=1, nmax=nsteps, exit_rules=True)
    ens.set_all(ic["xyz"], ic["velo"], ic["velof"], ic["eimp"], ic["tadd"])
    assert ens.run_md() == n * nsteps
    out = [ens.result(k) for k in range(n)]
    ens.close()
    return out


@pytest.mark.parametrize("name,n,nsteps", [("chloroethanol", 4, 20), ("caffeine", 3, 6)])
def test_md_with_refinement_matches_oracle_and_jacobi(qx, oracle, name, n, nsteps):
    from qcxms_b200 import ensemble_setup as es
    num, xyz, _ = qx.load_molecule(name)
    ic = es.synthetic_initial_conditions(num, xyz, n)
    with _Env(QCXMS_B200_CTA="576"):
        oa = _md(qx, num, ic, n, nsteps)          # seeds from the previous MD step for the first cycles, previous cycle afterwards
    with _Env(QCXMS_B200_CTA="576", QCXMS_B200_OA="0"):
        jac = _md(qx, num, ic, n, nsteps)
    for k in range(n):
        ref = oracle.md(num, ic["mass"], ic["xyz"][k], ic["velo"][k], ic["velof"][k], ic["eimp"][k], ic["tadd"][k], mchrg=1, nmax=nsteps)
        for got in (oa[k], jac[k]):
            assert got["nstep"] == ref["nstep"] == nsteps and got["scc_iter_total"] == ref["scc_iter_total"]
            assert np.array_equal(got["list"], ref["list"])
            assert np.abs(got["xyz"] - ref["xyz"]).max() < 1e-7 and np.abs(got["grad"] - ref["grad"]).max() < 2e-6
            assert abs(got["Epot"] - ref["Epot"]) < 1e-7 and np.abs(got["achrg"] - ref["achrg"]).max() < 1e-5
        assert np.abs(oa[k]["xyz"] - jac[k]["xyz"]).max() < 1e-9 and abs(oa[k]["Epot"] - jac[k]["Epot"]) < 1e-9


def test_single_points_on_wide_ctas(qx, oracle):
    """A lone single point (the level-1 entry point, one call per MD step from a Fortran host) runs on a wide CTA: previous-cycle seeds only."""
    num, xyz, _ = qx.load_molecule("caffeine")
    rng = np.random.default_rng(3)
    for k in range(4):
        x = xyz + 0.08 * rng.standard_normal(xyz.shape)
        q, e, g, stat = qx.get_xtb_egrad(num, x, 1, 2, qx.gfn2_xtb, 5000.0)
        ref = oracle.egrad(num, x, charge=1, multiplicity=2, etemp=5000.0)
        assert stat == 0 and abs(e - ref["energy"]) < 1e-8 and np.abs(g - ref["gradient"]).max() < 1e-6 and np.abs(q - ref["qat"]).max() < 1e-6


@pytest.mark.parametrize("name,n,nsteps,cta", [("caffeine", 200, 4, "320"), ("dichlorobenzamide_h", 160, 6, "320"), ("caffeine", 3, 5, "576")])
def test_last_sweep_as_dmma_products_changes_nothing(qx, oracle, name, n, nsteps, cta):
    """jacobi_polish (csrc/qx_device.cuh): the last Jacobi sweep of every eigen-decomposition runs as a Gram product plus one or two
    correction products.  Same SCC cycle counts, energies / positions to 1e-9 against the classical sweeps (QCXMS_B200_POLISH=0),
    and the usual gates against the oracle."""
    from qcxms_b200 import ensemble_setup as es
    num, xyz, _ = qx.load_molecule(name)
    ic = es.synthetic_initial_conditions(num, xyz, n)
    with _Env(QCXMS_B200_CTA=cta, QCXMS_B200_OA="0"):
        new = _md(qx, num, ic, n, nsteps)
    with _Env(QCXMS_B200_CTA=cta, QCXMS_B200_OA="0", QCXMS_B200_POLISH="0"):
        old = _md(qx, num, ic, n, nsteps)
    for a, b in zip(new, old):
        assert a["scc_iter_total"] == b["scc_iter_total"] and np.array_equal(a["list"], b["list"])
        assert np.abs(a["xyz"] - b["xyz"]).max() < 1e-9 and abs(a["Epot"] - b["Epot"]) < 1e-9
        assert np.abs(a["achrg"] - b["achrg"]).max() < 1e-8
    ref = oracle.md(num, ic["mass"], ic["xyz"][0], ic["velo"][0], ic["velof"][0], ic["eimp"][0], ic["tadd"][0], mchrg=1, nmax=nsteps)
    assert new[0]["scc_iter_total"] == ref["scc_iter_total"] and abs(new[0]["Epot"] - ref["Epot"]) < 1e-7
    assert np.abs(new[0]["xyz"] - ref["xyz"]).max() < 1e-7


def test_last_sweep_as_dmma_products_large_basis(qx, oracle):
    """jacobi_polish_gen on the global-slab path (C32H66, 194 AOs) against the classical sweeps and the oracle."""
    num, xyz, _ = qx.load_molecule("alkane_c32")
    rng = np.random.default_rng(11)
    geoms = xyz[None] + 0.03 * rng.standard_normal((3,) + xyz.shape)
    new = qx.egrad_batch(num, geoms, 1, 2, qx.gfn2_xtb, 5000.0)
    with _Env(QCXMS_B200_POLISH="0"):
        old = qx.egrad_batch(num, geoms, 1, 2, qx.gfn2_xtb, 5000.0)
    assert (new["stat"] == 0).all() and np.array_equal(new["niter"], old["niter"])
    assert np.abs(new["energy"] - old["energy"]).max() < 1e-9 and np.abs(new["gradient"] - old["gradient"]).max() < 1e-8
    ref = oracle.egrad(num, geoms[0], charge=1, multiplicity=2, etemp=5000.0)
    assert abs(new["energy"][0] - ref["energy"]) < 1e-8 and np.abs(new["gradient"][0] - ref["gradient"]).max() < 1e-6
