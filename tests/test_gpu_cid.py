"""GPU parity of the CID collision MD (reference src/cid.f90) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

from qcxms_b200 import ensemble_setup as es

pytestmark = pytest.mark.gpu


def _compare(got, ref, k):
    for key in ("status", "stopcid", "nstep", "nfrag", "collided", "scc_iter_total"):
        assert got[key][k] == ref[key], key
    assert np.array_equal(got["list"][k], ref["list"])
    assert np.abs(got["xyz"][k] - ref["xyz"]).max() < 1e-7
    assert np.abs(got["velo"][k] - ref["velo"]).max() < 1e-9
    assert np.abs(got["grad"][k] - ref["grad"]).max() < 2e-6
    assert np.abs(got["achrg"][k] - ref["achrg"]).max() < 1e-5
    assert np.abs(got["axyz"][k] - ref["axyz"]).max() < 1e-7
    assert np.abs(got["direc"][k] - ref["direc"]).max() < 1e-12
    assert abs(got["velo_cm"][k] - ref["velo_cm"]) < 1e-3 * max(1.0, abs(ref["velo_cm"]))   # m/s
    assert abs(got["aTlast"][k] - ref["aTlast"]) < 1e-3 and abs(got["ttime"][k] - ref["ttime"]) < 1e-9
    assert abs(got["epot"][k] - ref["epot"]) < 1e-7


def test_cid_first_and_second_collision_match_oracle(qx, oracle):
    num, xyz, _ = qx.load_molecule("chloroethanol")
    nt = 4
    ic = es.synthetic_initial_conditions(num, xyz, nt, first_id=20)
    mass = ic["mass"]
    rng = np.random.default_rng(99)
    rnd = rng.random((nt, 9))
    cfg = qx.cid_config(mchrg=1, gas="ar", elab=40.0, ntot=23)
    got = qx.cid(cfg, num, mass, 1, ic["xyz"], ic["velo"], rnd)
    refs = [oracle.cid(cfg, num, mass, 1, ic["xyz"][k], ic["velo"][k], rnd[k]) for k in range(nt)]
    for k in range(nt):
        _compare(got, refs[k], k)
    assert np.all(got["nstep"] == 23)
    # second collision continues from the output of the first (velocities kept, gas atom placed along direc)
    rnd2 = rng.random((nt, 9))
    cfg2 = qx.cid_config(mchrg=1, gas="ar", elab=40.0, ntot=12)
    got2 = qx.cid(cfg2, num, mass, 2, got["xyz"], got["velo"], rnd2, velo_cm=got["velo_cm"], direc=got["direc"], collided=got["collided"])
    for k in range(nt):
        ref2 = oracle.cid(cfg2, num, mass, 2, got["xyz"][k], got["velo"][k], rnd2[k], velo_cm=got["velo_cm"][k], direc=got["direc"][k],
                          collided=int(got["collided"][k]))
        _compare(got2, ref2, k)


def test_cid_head_on_collision_is_detected(qx, oracle):
    """Start the gas atom close (manual step distance) so that approach, turning point and the 5 consecutive
    'moving away' checks all happen within a short run: collided flips and total_steps is re-armed."""
    num, xyz, _ = qx.load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 2, first_id=40)
    rnd = np.tile(np.array([0.1, 0.2, 0.3, 0.7, 0.25, 0.0, 0.0, 0.3, 0.3]), (2, 1))   # f = g = 0: no lateral offset
    cfg = qx.cid_config(mchrg=1, gas="ar", elab=60.0, ntot=400, eexact=True, manual_dist=1)
    got = qx.cid(cfg, num, ic["mass"], 1, ic["xyz"], ic["velo"], rnd)
    ref = oracle.cid(cfg, num, ic["mass"], 1, ic["xyz"][0], ic["velo"][0], rnd[0])
    assert got["status"][0] == ref["status"] == 1
    assert ref["collided"] == 1 and ref["nfrag"] == 2 and ref["nstep"] < 400     # the scenario does what it is meant to
    # distance checks happen every 10 steps: allow one check of slack for SCC-threshold noise amplified by the impact
    assert got["collided"][0] == ref["collided"] and abs(got["nstep"][0] - ref["nstep"]) <= 10 and got["nfrag"][0] == ref["nfrag"]
    assert np.array_equal(got["list"][0], ref["list"])


def test_cid_n2_collision_gas_matches_oracle(qx, oracle):
    """N2 as collision gas: two nitrogen atoms (the second 1.09 A above the first), distance checks on the last one
    (reference src/cid.f90:179-180, 660-667, 1124-1139)."""
    num, xyz, _ = qx.load_molecule("chloroethanol")
    nt = 3
    ic = es.synthetic_initial_conditions(num, xyz, nt, first_id=60)
    rng = np.random.default_rng(5)
    rnd = rng.random((nt, 9))
    cfg = qx.cid_config(mchrg=1, gas="n2", elab=40.0, ntot=15)
    got = qx.cid(cfg, num, ic["mass"], 1, ic["xyz"], ic["velo"], rnd)
    for k in range(nt):
        _compare(got, oracle.cid(cfg, num, ic["mass"], 1, ic["xyz"][k], ic["velo"][k], rnd[k]), k)
    assert np.all(got["nstep"] == 15) and np.all(got["status"] == 1)
