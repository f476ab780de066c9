#!/usr/bin/env python
"""Production-like soak run (development aid): caffeine EI ensemble with the exit rules on, synthetic initial conditions.
usage: python tools/soak_production.py [ntraj] [nmax]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcxms_b200 as qx
from qcxms_b200 import ensemble_setup as es
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 592
nmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
num, xyz, _ = qx.load_molecule("caffeine")
ic = es.synthetic_initial_conditions(num, xyz, nt, first_id=0)
ens = qx.Ensemble(num, ic["mass"], nt, mchrg=1, nmax=nmax, exit_rules=True)
ens.set_all(ic["xyz"], ic["velo"], ic["velof"], ic["eimp"], ic["tadd"])
t0 = time.perf_counter()
steps = ens.run_md()
dt = time.perf_counter() - t0
r = ens.results()
bins, _ = ens.histogram(256)
tm = ens.last_timing()
ens.close()
print("ntraj %d nmax %d: %d trajectory-steps in %.2f s wall (%.0f steps/s; kernel %.2f s, %d launches), SCC cycles/step %.2f" %
      (nt, nmax, steps, dt, steps / dt, tm["kernel_ms"] / 1e3, tm["launches"], tm["scc_iterations"] / max(steps, 1)))
print("status: finished %d, failed %d, running %d; mdok %d; nfrag histogram %s; mean steps %.0f" %
      ((r["status"] == 1).sum(), (r["status"] == 2).sum(), (r["status"] == 0).sum(), r["mdok"].sum(), np.bincount(r["nfrag"], minlength=5).tolist(), r["nstep"].mean()))
print("largest peaks (m/z: count):", {int(i): int(bins[i]) for i in np.argsort(bins)[::-1][:8] if bins[i] > 0})
