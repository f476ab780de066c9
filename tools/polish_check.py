#!/usr/bin/env python
"""Development aid: the Gram-matrix polish (jacobi_polish) against classical sweeps on distorted cations.
    QCXMS_B200_POLISH=0 python tools/polish_check.py a.npz; python tools/polish_check.py b.npz; python tools/polish_check.py a.npz b.npz"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 3:
    a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
    for k in ("energy", "gradient", "qat"):
        print("%-8s max |diff| %.3e" % (k, np.abs(a[k] - b[k]).max()))
    print("niter equal:", bool((a["niter"] == b["niter"]).all()), " stat ok:", bool((a["stat"] == 0).all() and (b["stat"] == 0).all()),
          " time %.3f vs %.3f s" % (float(a["dt"]), float(b["dt"])))
    sys.exit(0)
import qcxms_b200 as qx
rng = np.random.default_rng(0)
res = {}
dt = 0.0
CASES = [("caffeine", 592), ("dichlorobenzamide_h", 300), ("thf_h", 300), ("chloroethanol", 300)]
if os.environ.get("QX_POLISH_BIG"):
    CASES = [("alkane_c14", 296), ("alkane_c17", 296), ("alkane_c32", 148), ("peptide_cl", 148)]
for name, nsys in CASES:
    num, xyz, _ = qx.load_molecule(name)
    geoms = xyz[None] + 0.05 * rng.standard_normal((nsys,) + xyz.shape)
    qx.egrad_batch(num, geoms[:4], 1, 2, qx.gfn2_xtb, 5000.0)
    t0 = time.perf_counter()
    out = qx.egrad_batch(num, geoms, 1, 2, qx.gfn2_xtb, 5000.0)
    t = time.perf_counter() - t0
    dt += t
    print("%s: %.1f egrad/s, SCC cycles %.2f, bad %d" % (name, nsys / t, out["niter"].mean(), int((out["stat"] != 0).sum())), flush=True)
    for k in ("energy", "gradient", "qat", "niter", "stat"):
        res.setdefault(k, []).append(np.asarray(out[k], dtype=np.float64).ravel())
np.savez(sys.argv[1], dt=dt, **{k: np.concatenate(v) for k, v in res.items()})
