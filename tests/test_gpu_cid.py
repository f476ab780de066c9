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


def _mfp_compare(got, ref, k, nv):
    for key in ("status", "mdok", "fragstate", "nstep", "nfrag", "scc_iter_total"):
        assert got[key][k] == ref[key], key
    assert np.array_equal(got["list"][k], ref["list"])
    assert np.abs(got["xyz"][k] - ref["xyz"]).max() < 1e-7
    assert np.abs(got["velo"][k] - ref["velo"]).max() < 1e-9
    assert np.abs(got["grad"][k] - ref["grad"]).max() < 2e-6
    assert np.abs(got["achrg"][k] - ref["achrg"]).max() < 1e-5
    assert np.abs(got["axyz"][k] - ref["axyz"]).max() < 1e-7
    for key, tol in (("Tav", 1e-3), ("aTlast", 1e-3), ("Epav", 1e-7), ("Ekav", 1e-7), ("Epot", 1e-7), ("Ekin", 1e-7), ("ttime", 1e-9), ("dtime", 1e-9)):
        assert abs(got[key][k] - ref[key]) < tol, key
    assert abs(nv[k] - ref["new_velo"]) < 1e-3 * max(1.0, abs(ref["new_velo"]))       # m/s


def test_mfp_md_matches_oracle(qx, oracle):
    """Mean-free-path md() of a CID run (reference md() with method 3, icoll >= 1; src/md.f90:246-255, 466-621): drifting ions, one
    of them losing its chlorine so that the fragment averaging and the moved end of the run are exercised."""
    num, xyz, _ = qx.load_molecule("chloroethanol")
    nt = 4
    ic = es.synthetic_initial_conditions(num, xyz, nt, first_id=70)
    mass = ic["mass"]
    velo = ic["velo"].copy()
    velo[:, :, 2] += 9000.0 / 2.18769126364e+06                 # 9 km/s drift along z
    icl = int(np.nonzero(num == 17)[0][0])
    x0 = ic["xyz"][3]
    ic_c = int(np.argmin(np.where(num == 6, np.linalg.norm(x0 - x0[icl], axis=1), 1e9)))
    velo[3, icl] += 2.5e-3 * (x0[icl] - x0[ic_c]) / np.linalg.norm(x0[icl] - x0[ic_c])
    nv_in = np.array([9000.0, 8990.0, 9010.0, 9000.0])
    ens = qx.Ensemble(num, mass, nt, mchrg=1, nmax=60, isec=2)
    ens.set_all(ic["xyz"], velo, np.ones((nt, len(num))), np.zeros(nt), np.zeros(nt))
    ens.set_mfp(2, nv_in)
    ens.run_md()
    got = ens.results()
    nv = ens.new_velo()
    ens.close()
    refs = [oracle.md_mfp(num, mass, ic["xyz"][k], velo[k], 2, nv_in[k], nmax=60, isec=2) for k in range(nt)]
    for k in range(nt):
        _mfp_compare(got, refs[k], k, nv)
    assert np.all(got["nstep"][:3] == 60) and refs[3]["nfrag"] == 2 and refs[3]["nstep"] < 60


def test_mfp_md_averages_fragment_structures(qx, oracle):
    """14 atoms: add_steps = 500 after the fragmentation, the structure is averaged over the 50 steps that follow
    (src/md.f90:233-235, 504-621); bounded at 90 steps through the step limit of both implementations."""
    num, xyz, _ = qx.load_molecule("thf_h")
    nt = 2
    ic = es.synthetic_initial_conditions(num, xyz, nt, first_id=80)
    mass = ic["mass"]
    velo = ic["velo"].copy()
    ih = int(np.nonzero(num == 1)[0][0])
    for k in range(nt):
        x0 = ic["xyz"][k]
        heavy = int(np.argmin(np.where(num > 1, np.linalg.norm(x0 - x0[ih], axis=1), 1e9)))
        velo[k, ih] += (0.02 + 0.004 * k) * (x0[ih] - x0[heavy]) / np.linalg.norm(x0[ih] - x0[heavy])
    nv_in = np.zeros(nt)
    ens = qx.Ensemble(num, mass, nt, mchrg=1, nmax=40, isec=2)
    ens.set_all(ic["xyz"], velo, np.ones((nt, len(num))), np.zeros(nt), np.zeros(nt))
    ens.set_mfp(1, nv_in)
    ens.run_md(max_steps=90)
    got = ens.results()
    nv = ens.new_velo()
    ens.close()
    for k in range(nt):
        ref = oracle.md_mfp(num, mass, ic["xyz"][k], velo[k], 1, 0.0, nmax=40, isec=2, max_steps=90)
        assert ref["nfrag"] == 2 and ref["nstep"] == 90          # ran past nmax = 40 because the fragmentation moved the end
        got["status"][k] = ref["status"]                         # paused by the step limit: the oracle reports 1, the ensemble "running"
        _mfp_compare(got, ref, k, nv)
        assert np.abs(ref["axyz"] - ref["xyz"]).max() > 1e-3     # axyz is the 50-step average, not the last structure


def test_cid_production_run_matches_oracle_backend(qx, oracle):
    """The collision loop of a CID run (main.F90:1490-2163) with the CUDA back ends against the same driver on the CPU oracle."""
    from qcxms_b200 import production as prod
    num, xyz, _ = qx.load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 3, first_id=90)
    kw = dict(mchrg=1, gas="ar", elab=40.0, run_type="maxcoll", max_coll=2, minmass=20, first_itrj=1, seed=5, cid_ntot=20, mfp_nmax=15)
    got = prod.run_cid(num, ic["mass"], ic["xyz"], ic["velo"], **kw)
    ref = prod.run_cid(num, ic["mass"], ic["xyz"], ic["velo"], cid_batch=oracle.cid_batch, mfp_batch=oracle.mfp_batch, energies=oracle.energies, **kw)
    assert got["records"] == ref["records"] and len(got["records"]) == 3
    for a, b in zip(got["per_traj"], ref["per_traj"]):
        assert [(e["kind"], e["icoll"], e["nstep"], e["nfrag"]) for e in a["events"]] == [(e["kind"], e["icoll"], e["nstep"], e["nfrag"]) for e in b["events"]]
        va = [e.get("velo_cm", e.get("new_velo")) for e in a["events"]]
        vb = [e.get("velo_cm", e.get("new_velo")) for e in b["events"]]
        assert np.allclose(va, vb, rtol=1e-6, atol=1e-3)


def test_esi_heating_md_matches_oracle(qx, oracle):
    """md() as the heating MD before the first collision of an ESI/CID run (reference md() with method 3, icoll = 0, starting_md;
    src/md.f90:428-434, call site main.F90:1357-1362): Berendsen scaling towards tscale during the first nadd steps."""
    num, xyz, _ = qx.load_molecule("thf_h")
    nt = 3
    ic = es.synthetic_initial_conditions(num, xyz, nt, first_id=40, temperature=300.0)
    mass, nmax = ic["mass"], 40
    tscale = 900.0
    e_scale = 1.5 * len(num) * 3.166808578545117e-06 * tscale          # inner energy for tscale (Eh)
    pretadd = nmax * 41.3413733365614 * 0.5 * 0.75                      # main.F90:1336 with tstep 0.5 fs
    ens = qx.Ensemble(num, mass, nt, mchrg=1, nmax=nmax, isec=1)
    ens.set_all(ic["xyz"], ic["velo"], np.ones((nt, len(num))), np.full(nt, e_scale), np.full(nt, pretadd))
    ens.set_esi(tscale)
    assert ens.run_md() == nt * nmax
    got = ens.results()
    ens.close()
    for k in range(nt):
        ref = oracle.md_esi(num, mass, ic["xyz"][k], ic["velo"][k], tscale, e_scale, pretadd, nmax=nmax, isec=1)
        assert got["nstep"][k] == ref["nstep"] == nmax and got["mdok"][k] == ref["mdok"] == 1 and got["scc_iter_total"][k] == ref["scc_iter_total"]
        assert np.abs(got["xyz"][k] - ref["xyz"]).max() < 1e-7 and np.abs(got["velo"][k] - ref["velo"]).max() < 1e-9
        assert abs(got["Epot"][k] - ref["Epot"]) < 1e-7 and abs(got["Ekin"][k] - ref["Ekin"]) < 1e-9
        assert abs(got["aTlast"][k] - ref["aTlast"]) < 1e-3 and abs(got["Tav"][k] - ref["Tav"]) < 1e-3
        # the thermostat heats: the last temperatures are well above the 300 K start
        assert ref["aTlast"] > 350.0


@pytest.mark.parametrize("name,elab", [("thf_h", 40.0), ("dichlorobenzamide_h", 60.0)])
def test_cid_on_the_reference_cid_examples(qx, oracle, name, elab):
    """The reference's own CID inputs (share/examples/CID: protonated THF, elab 40; protonated dichlorobenzamide, elab 60; Ar): the first
    collision against the oracle."""
    num, xyz, _ = qx.load_molecule(name)
    nt = 3
    ic = es.synthetic_initial_conditions(num, xyz, nt, first_id=5, temperature=300.0)
    rnd = np.random.default_rng(17).random((nt, 9))
    cfg = qx.cid_config(mchrg=1, gas="ar", elab=elab, ntot=14)
    got = qx.cid(cfg, num, ic["mass"], 1, ic["xyz"], ic["velo"], rnd)
    for k in range(nt):
        _compare(got, oracle.cid(cfg, num, ic["mass"], 1, ic["xyz"][k], ic["velo"][k], rnd[k]), k)
