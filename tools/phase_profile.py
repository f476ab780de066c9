#!/usr/bin/env python
"""Per-phase cycle breakdown of the fused egrad kernel (profiling build with -DQX_PROFILE_PHASES).
Usage (on the GPU box): python tools/phase_profile.py [molecule] [nsys]
Builds qcxms_b200/libqcxms_b200_prof.so if needed; never used by the product path."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, "qcxms_b200", "libqcxms_b200_prof.so")
NAMES = ["cn+rep", "d4 nonsc/ATM", "coulomb+integrals", "cholesky basis", "broyden", "potential", "build H1", "transform C^T H C",
         "jacobi", "fermi", "density", "mulliken", "scc energy", "W matrix", "grad AO pairs", "grad rest"]
def build():
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-DQX_PROFILE_PHASES",
                           "-Xcompiler", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "qcxms_b200", "csrc", "cabi.cu")])
if __name__ == "__main__":
    if "--build" in sys.argv:
        build(); sys.exit(0)
    import qcxms_b200.api as api
    api.LIB_PATH = so
    mol = sys.argv[1] if len(sys.argv) > 1 else "caffeine"
    nsys = int(sys.argv[2]) if len(sys.argv) > 2 else 592
    num, xyz, _ = api.load_molecule(mol)
    rng = np.random.default_rng(0)
    geoms = xyz[None] + 0.05 * rng.standard_normal((nsys,) + xyz.shape)
    api.egrad_batch(num, geoms[:8], 1, 2, 2, 5000.0)
    buf = (C.c_double * 16)()
    api.lib().qcxms_b200_debug_phase_cycles(buf)
    import time
    t0 = time.perf_counter()
    out = api.egrad_batch(num, geoms, 1, 2, 2, 5000.0)
    dt = time.perf_counter() - t0
    api.lib().qcxms_b200_debug_phase_cycles(buf)
    cyc = np.array(list(buf)); tot = cyc.sum()
    print("%s: %d systems in %.3f s (%.1f egrad/s), mean SCC cycles %.2f" % (mol, nsys, dt, nsys / dt, out["niter"].mean()))
    for n, c in zip(NAMES, cyc):
        print("  %-20s %6.2f %%   %10.0f cycles/egrad" % (n, 100 * c / tot, c / nsys))
