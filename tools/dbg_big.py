import sys; sys.path.insert(0, '.')
import numpy as np, time
import qcxms_b200 as qx
from oracle import pyoracle as po
num, xyz, _ = qx.load_molecule("alkane_c32")
rng = np.random.default_rng(2)
x = xyz + 0.03 * rng.standard_normal(xyz.shape)
t = time.time()
out = qx.egrad_batch(num, x[None], 1, 2, qx.gfn2_xtb, 5000.0)
print("gpu time", time.time() - t, "stat", out["stat"], "niter", out["niter"], "E", out["energy"])
ref = po.egrad(num, x, 1, 2, 2, 5000.0, detail=True)
print("ref", ref["stat"], ref["niter"], ref["energy"], "dE", out["energy"][0] - ref["energy"], "dG", abs(out["gradient"][0] - ref["gradient"]).max(), "dq", abs(out["qat"][0] - ref["qat"]).max())
