"""Ensemble set-up of an EI run: IEE distribution, MO mapping, heating times on the host; ground-state md() modes
(it = -1 / 0) on the GPU against the oracle."""
import numpy as np
import pytest

from qcxms_b200 import setup_ei as se
from qcxms_b200.api import load_molecule


def test_iee_distribution_parameters():
    # caffeine-like numbers: 74 valence electrons, 24 atoms, 70 eV electrons, HOMO at -0.40 Eh
    ieeel, exc, nuc = 74.0, (70.0 / 27.21138505 + 0.40) * 27.21138505, 24
    a, b = se.getieeab(ieeel, 1, exc, nuc, 0.6)
    ieemax, pmax, eavg = se.getmaxiee(a, b, ieeel, 1, exc)
    assert 0.0 < a <= 0.3 + 1e-7 and b > 0.0
    assert eavg / nuc >= 0.6 and eavg / nuc < 0.75            # the first (a, b) on the 0.005 / 0.035 grid that reaches ieeatm per atom
    assert 0 < ieemax < exc and pmax > 0
    xs = np.arange(0.5, exc, 0.5)
    assert np.all(se.poiss0(a, b, ieeel, xs) <= pmax * (1 + 1e-9))
    # the previous grid point does not reach the average yet
    _, _, eprev = se.getmaxiee(min(a - 0.005, 0.3), b - 0.035, ieeel, 1, exc)
    assert eprev / nuc < 0.6
    # Gaussian variant: maximum at ieeel * iee_b
    assert abs(se.gauss0(0.2, 0.5, 10.0, 5.0) - 1.0) < 1e-15 and se.gauss0(0.2, 0.5, 10.0, 9.0) < 1.0


def test_momap_and_calctrelax():
    rng = np.random.default_rng(3)
    emo = np.sort(rng.uniform(-1.0, 0.5, 40))
    ihomo = 20
    target = emo[4] + emo[17]
    best = np.inf
    mo1, mo2 = se.momap(ihomo, emo, target, np.random.default_rng(5))
    assert 1 <= mo1 <= ihomo and 0 <= mo2 <= ihomo // 2
    # the result is the best of the 5001 random trials it drew: reproduce them
    r2 = np.random.default_rng(5)
    for _ in range(5001):
        a, b = se.irand(ihomo, r2), se.irand(ihomo, r2)
        v = se.irand(ihomo // 2, r2) + ihomo
        d = emo[a - 1] + (0.0 if b > ihomo // 2 else emo[b - 1]) + emo[v - 1]
        best = min(best, abs(d - target))
    d = emo[mo1 - 1] + (emo[mo2 - 1] if mo2 > 0 else 0.0)
    assert min(abs(d + emo[v - 1] - target) for v in range(ihomo + 1, ihomo + ihomo // 2 + 1)) <= best + 1e-15
    # calctrelax: a sum of trelax * exp(alp * (e_k - e_j)) over the occupied MOs above i
    t = se.calctrelax(emo, ihomo, 18, 2000.0)
    ref = 2000.0 * (np.exp(0.5 * 27.21138505 * (emo[17] - emo[18])) + np.exp(0.5 * 27.21138505 * (emo[18] - emo[19])))
    assert abs(t - ref) < 1e-9 * ref and se.calctrelax(emo, ihomo, ihomo, 2000.0) == 0.0
    assert all(1 <= se.irand(7, rng) <= 7 for _ in range(200))


def test_mdinitu_energy():
    num, _, _ = load_molecule("caffeine")
    mass = np.array([{1: 1.008, 6: 12.011, 7: 14.007, 8: 15.999}[int(z)] for z in num]) * se.AMUTOAU
    e = 3.0 * 0.5 * se.KB * 500.0 * len(num)
    v = se.mdinitu(mass, e, np.random.default_rng(0))
    assert abs(0.5 * (mass[:, None] * v * v).sum() - e) < 1e-12     # every component carries e / (3 nat)


@pytest.mark.gpu
@pytest.mark.parametrize("it", [-1, 0])
def test_ground_state_md_modes_match_oracle(qx, oracle, it):
    num, xyz, _ = qx.load_molecule("chloroethanol")
    mass = np.array([qx.api.AMUTOAU * {1: 1.00794075, 6: 12.0107359, 8: 15.99940492, 17: 35.45293758}[int(z)] for z in num])
    velo = se.mdinitu(mass, 3.0 * 0.5 * se.KB * 500.0 * len(num), np.random.default_rng(2))
    if it < 0:
        velo = velo * 1.4        # twice the target temperature: the thermostat has to act after step 50
    nmax = 70 if it < 0 else 30
    ens = qx.Ensemble(num, mass, 1, mchrg=0, nmax=nmax, etemp=298.15)
    ens.set_all(xyz[None], velo[None], np.ones((1, len(num))), np.zeros(1), np.zeros(1))
    ens.set_gs_mode(it, 500.0)
    assert ens.run_md() == nmax
    got = ens.result(0)
    ref = oracle.md_gs(num, mass, xyz, velo, it, 500.0, 298.15, mchrg=0, nmax=nmax)
    assert got["nstep"] == ref["nstep"] == nmax and got["mdok"] == ref["mdok"] == 1 and got["scc_iter_total"] == ref["scc_iter_total"]
    assert np.abs(got["xyz"] - ref["xyz"]).max() < 1e-7 and np.abs(got["velo"] - ref["velo"]).max() < 1e-9
    assert abs(got["Tav"] - ref["Tav"]) < 1e-3 and abs(got["Epav"] - ref["Epav"]) < 1e-7
    if it == 0:
        rec = ens.gs_records(0)
        assert rec.shape == (nmax, len(num), 6) and np.abs(rec - ref["gs"]).max() < 1e-7
        assert np.array_equal(rec[0][:, :3], xyz) and np.array_equal(rec[0][:, 3:], velo)      # the first record is the start point
    else:
        # the run was rescaled: the final temperature moved towards the target
        t_end = (mass[:, None] * got["velo"] ** 2).sum() / (3 * len(num) * se.KB)
        t_start = (mass[:, None] * velo ** 2).sum() / (3 * len(num) * se.KB)
        assert t_end < 0.9 * t_start
    ens.close()


@pytest.mark.gpu
def test_setup_to_start_directories(qx, tmp_path):
    """neutral geometry -> ground-state MD -> IEE draw -> TMPQCXMS/TMP.n -> production md() from the files"""
    from qcxms_b200 import startfiles as sf
    num, xyz, _ = qx.load_molecule("chloroethanol")
    mass = np.array([qx.api.AMUTOAU * {1: 1.00794075, 6: 12.0107359, 8: 15.99940492, 17: 35.45293758}[int(z)] for z in num])
    rng = np.random.default_rng(11)
    gs = se.ground_state_sampling(num, mass, xyz, nmax0=40, rng=rng)
    assert gs["mdok"] == 1 and gs["records"].shape == (40, len(num), 6) and 100.0 < gs["Tav"] < 1500.0
    runs = se.prepare_runs(num, mass, gs["records"], ntraj=6, rng=rng)
    assert len(runs["eimp"]) == 6 and np.all(runs["eimp"] > 0) and np.all(runs["tadd"] >= 200.0 * qx.api.FSTOAU * 0.999)
    assert np.all(runs["velof"] <= 1.0 + 1e-12) and np.all(runs["velof"].max(axis=1) == 1.0)
    assert len(set(runs["step"])) == 6                                   # distinct snapshots of the ground-state trajectory
    se.write_directories(str(tmp_path), num, runs)
    st = sf.read_start(str(tmp_path / "TMPQCXMS" / "TMP.3"))
    assert st["itrj"] == 3 and abs(st["eimp"] - runs["eimp"][2]) < 1e-13 and np.abs(st["velof"] - runs["velof"][2]).max() < 1e-13
    ens = qx.Ensemble(num, mass, 6, mchrg=1, nmax=10)
    ens.set_all(runs["xyz"], runs["velo"], runs["velof"], runs["eimp"], runs["tadd"])
    assert ens.run_md() == 60
    ens.close()
