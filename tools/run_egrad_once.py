#!/usr/bin/env python
"""Driver for profiler captures: one batched egrad launch (warm-up launch first)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcxms_b200 as qx
mol = sys.argv[1] if len(sys.argv) > 1 else "caffeine"
nsys = int(sys.argv[2]) if len(sys.argv) > 2 else 592
num, xyz, _ = qx.load_molecule(mol)
rng = np.random.default_rng(0)
geoms = xyz[None] + 0.05 * rng.standard_normal((nsys,) + xyz.shape)
qx.egrad_batch(num, geoms[:8], 1, 2, qx.gfn2_xtb, 5000.0)
out = qx.egrad_batch(num, geoms, 1, 2, qx.gfn2_xtb, 5000.0)
print("niter mean", out["niter"].mean(), "stat", int(np.abs(out["stat"]).sum()))
