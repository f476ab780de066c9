#!/usr/bin/env python
"""CID production soak run (development aid; BASELINE config 4): the reference's share/examples/CID/Tetrahydrofuran input
(`xtb2 / cid / elab 40 / maxcoll 6 / noesi`, Ar) for an ensemble of protonated THF ions with synthetic 500 K initial conditions,
through qcxms_b200.production.run_cid (cid() and the mean-free-path md() on the GPU, fragment single points batched).
usage: python tools/soak_cid.py [ntraj] [molecule] [elab] [max_coll] [out.json]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcxms_b200 as qx
from qcxms_b200 import ensemble_setup as es
from qcxms_b200 import production as prod
from qcxms_b200.fragments import spectrum_from_records

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mol = sys.argv[2] if len(sys.argv) > 2 else "thf_h"
elab = float(sys.argv[3]) if len(sys.argv) > 3 else 40.0
max_coll = int(sys.argv[4]) if len(sys.argv) > 4 else 6
out = sys.argv[5] if len(sys.argv) > 5 else None
num, xyz, chg = qx.load_molecule(mol)
ic = es.synthetic_initial_conditions(num, xyz, nt, first_id=0)
t0 = time.perf_counter()
res = prod.run_cid(num, ic["mass"], ic["xyz"], ic["velo"], mchrg=1, gas="ar", elab=elab, run_type="maxcoll", max_coll=max_coll, minmass=20, seed=1)
dt = time.perf_counter() - t0
ev = [e for t in res["per_traj"] for e in t["events"]]
nfatal = sum(1 for e in ev if e["kind"] == "fatal")
steps_cid = sum(e["nstep"] for e in ev if e["kind"] == "cid")
steps_mfp = sum(e["nstep"] for e in ev if e["kind"] == "mfp")
ncoll = [max([e["icoll"] for e in t["events"]] or [0]) for t in res["per_traj"]]
nfragev = sum(1 for e in ev if e["nfrag"] > 1)
spec = spectrum_from_records(res["records"], 256)
peaks = {int(i): float(spec[i]) for i in np.argsort(spec)[::-1][:8] if spec[i] > 0}
summary = dict(molecule=mol, ntraj=nt, elab_eV=elab, max_coll=max_coll, wall_s=dt, cid_steps=int(steps_cid), mfp_steps=int(steps_mfp),
               steps_per_s=(steps_cid + steps_mfp) / dt, cid_calls=sum(1 for e in ev if e["kind"] == "cid"), mfp_calls=sum(1 for e in ev if e["kind"] == "mfp"),
               collisions_per_traj=float(np.mean(ncoll)), fragmenting_events=int(nfragev), stopcid=sum(int(e.get("stopcid", 0)) for e in ev),
               fatal_fragment_single_points=nfatal, records=len(res["records"]), charge_sum_per_traj=float(sum(float(r[:10]) for r in res["records"]) / nt), peaks=peaks)
print(json.dumps(summary))
if out:
    with open(out, "w") as f:
        json.dump(summary, f, indent=1)
