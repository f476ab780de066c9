"""EI production cascade (md -> manage_fragments -> secondary trajectories), host logic on the CPU oracle back end."""
import numpy as np

from qcxms_b200 import ensemble_setup as es
from qcxms_b200 import production as prod
from qcxms_b200.api import load_molecule


def _ic(nt, first_id):
    num, xyz, _ = load_molecule("chloroethanol")
    return num, es.synthetic_initial_conditions(num, xyz, nt, first_id=first_id, ieeatm=2.5, tadd_fs=40.0)


def test_cascade_on_oracle_backend(oracle):
    num, ic = _ic(3, 900)
    out = prod.run_ei(num, ic["mass"], ic["xyz"], ic["velo"], ic["velof"], ic["eimp"], ic["tadd"], mchrg=1, nmax=240, maxsec=3, first_itrj=11,
                      md_batch=oracle.md_batch, energies=oracle.energies)
    assert len(out["per_traj"]) == 3
    nsec = 0
    for t in out["per_traj"]:
        gens = t["generations"]
        assert gens[0]["isec"] == 1 and gens[0]["nat"] == len(num) and all(g["isec"] == i + 1 for i, g in enumerate(gens))
        assert all(g["md_ok"] for g in gens[:-1])
        nsec += len(gens) > 1
        # the atoms of a secondary trajectory are the atoms of the fragment that kept the charge
        for a, b in zip(gens[:-1], gens[1:]):
            assert b["nat"] < a["nat"] and a["tcont"] > 0
        # charge conservation over the records of one trajectory: the statistical charges sum to the ion's charge
        q = sum(float(r[:10]) for r in t["records"])
        if gens[-1]["md_ok"]:
            assert abs(q - 1.0) < 1e-5
        assert all(int(r[13:18]) == t["itrj"] for r in t["records"])
    assert nsec >= 1          # at least one trajectory fragmented and continued (otherwise the test does not test the cascade)
    assert len(out["records"]) >= 3
