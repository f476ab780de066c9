#!/usr/bin/env python
"""Wall-clock throughput of the batched single point for one molecule (development aid): python tools/time_egrad.py name nsys"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcxms_b200 as qx
mol = sys.argv[1] if len(sys.argv) > 1 else "caffeine"
nsys = int(sys.argv[2]) if len(sys.argv) > 2 else 592
num, xyz, chg = qx.load_molecule(mol)
rng = np.random.default_rng(0)
geoms = xyz[None] + 0.03 * rng.standard_normal((nsys,) + xyz.shape)
qx.egrad_batch(num, geoms[:4], 1, 2, qx.gfn2_xtb, 5000.0)
t0 = time.perf_counter()
out = qx.egrad_batch(num, geoms, 1, 2, qx.gfn2_xtb, 5000.0)
dt = time.perf_counter() - t0
print("%s: nat %d, %d systems in %.3f s -> %.1f egrad/s, SCC cycles %.2f, stat!=0: %d" % (mol, len(num), nsys, dt, nsys / dt, out["niter"].mean(), int((out["stat"] != 0).sum())))
