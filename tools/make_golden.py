#!/usr/bin/env python
"""Regenerate tests/golden/*.json from the CPU oracle (regression pins for oracle AND CUDA path).

These are NOT tblite outputs (tblite cannot be built offline): parity with the reference's
tblite stays unpinned.  tools/golden_driver.f90 is the generator to run wherever tblite is installable; its
output has the same JSON layout and can replace this file unchanged.
"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import pyoracle as po
from qcxms_b200.api import load_molecule
from qcxms_b200 import ensemble_setup as es

out = {"generator": "oracle/xtb_oracle.c via tools/make_golden.py", "cases": []}
for name, chg, mult, etemp in [("chloroethanol", 0, 1, 300.0), ("chloroethanol", 1, 2, 5000.0), ("monoethanolamine", 1, 2, 5000.0),
                               ("thf_h", 1, 1, 5000.0), ("dichlorobenzamide_h", 1, 1, 5000.0), ("caffeine", 1, 2, 5000.0)]:
    num, xyz, _ = load_molecule(name)
    r = po.egrad(num, xyz, charge=chg, multiplicity=mult, etemp=etemp, detail=True)
    out["cases"].append(dict(molecule=name, charge=chg, multiplicity=mult, etemp=etemp, energy=r["energy"], niter=r["niter"],
                             gradient=r["gradient"].tolist(), qat=r["qat"].tolist(),
                             terms={k: r[k] for k in ("e_rep", "e_disp_atm", "e_disp_sc", "e_el", "e_es2", "e_es3", "e_aes", "e_ts")}))
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "egrad_golden.json"), "w"), indent=1)

# MD-side vectors
num, xyz, _ = load_molecule("chloroethanol")
ic = es.synthetic_initial_conditions(num, xyz, 2)
md = {"cases": []}
for k in range(2):
    r = po.md(num, ic["mass"], ic["xyz"][k], ic["velo"][k], ic["velof"][k], ic["eimp"][k], ic["tadd"][k], mchrg=1, nmax=10)
    md["cases"].append(dict(traj=k, nstep=r["nstep"], scc_iter_total=r["scc_iter_total"], Epot=r["Epot"], Ekin=r["Ekin"], Tav=r["Tav"],
                            xyz=r["xyz"].tolist(), velo=r["velo"].tolist(), list=r["list"].tolist()))
json.dump(md, open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "md_golden.json"), "w"), indent=1)
print("golden files written")
