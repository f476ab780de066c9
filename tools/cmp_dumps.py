import numpy as np, sys
ref = np.load("gpurun_out/dump_lp8.npz")
for t in ("t1e7", "t3e6", "t3e5"):
    d = np.load("gpurun_out/dump_%s.npz" % t)
    print(t, "dE %.2e dG %.2e dQ %.2e  niter equal: %s (mean %.3f vs %.3f)" % (np.abs(d["energy"] - ref["energy"]).max(), np.abs(d["gradient"] - ref["gradient"]).max(),
          np.abs(d["qat"] - ref["qat"]).max(), np.array_equal(d["niter"], ref["niter"]), d["niter"].mean(), ref["niter"].mean()))
