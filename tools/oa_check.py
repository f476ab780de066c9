#!/usr/bin/env python
"""Development aid: eigenpair refinement (wide CTAs, QCXMS_B200_CTA=576) against the Jacobi path and the oracle on distorted caffeine
cations.  Usage: QCXMS_B200_CTA=576 [QCXMS_B200_OA=0] [QCXMS_B200_OA_STOP=x] python tools/oa_check.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import qcxms_b200 as qx
from oracle import pyoracle as po
num, xyz, _ = qx.load_molecule("caffeine")
rng = np.random.default_rng(7)
geoms = xyz[None] + 0.08 * rng.standard_normal((12,) + xyz.shape)
de, dg, dq, dn = [], [], [], []
for k in range(12):
    q, e, g, stat = qx.get_xtb_egrad(num, geoms[k], 1, 2, 2, 5000.0)     # single points run on the wide CTAs (nwork = 1)
    ref = po.egrad(num, geoms[k], charge=1, multiplicity=2, etemp=5000.0, detail=True)
    de.append(abs(e - ref["energy"])); dg.append(np.abs(g - ref["gradient"]).max()); dq.append(np.abs(q - ref["qat"]).max())
print("max dE %.2e dG %.2e dQ %.2e" % (max(de), max(dg), max(dq)), " per geometry dE:", " ".join("%.1e" % v for v in de))
