#!/usr/bin/env python
"""Per-phase cycle breakdown of the fused egrad kernel (profiling build with -DQX_PROFILE_PHASES).
Usage (on the GPU box): python tools/phase_profile.py [molecule] [nsys]
Build the profiling library first (here, no GPU needed): python tools/phase_profile.py --build  -> build/libqcxms_b200_prof.so;
never used by the product path."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, "build", "libqcxms_b200_prof.so")
NAMES = ["cn+rep", "d4 nonsc/ATM", "coulomb+integrals", "cholesky basis", "broyden", "potential", "build H1", "transform C^T H C",
         "jacobi", "fermi", "density", "mulliken", "scc energy", "W matrix", "grad AO pairs", "grad rest"]
def build(out=so, extra=()):
    tag = "prof" + "".join(c if c.isalnum() else "_" for c in "".join(extra))
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "qcxms_b200", "csrc"), "-j8", "-s", "OUT=" + os.path.abspath(out),
                           "OBJDIR=" + os.path.join(ROOT, "build", "obj_" + tag), "EXTRA=-DQX_PROFILE_PHASES " + " ".join(extra)])
if __name__ == "__main__":
    # --so PATH: use (or with --build: write) another profiling library; -D... flags are passed to nvcc; --dump FILE: save results
    args = [a for a in sys.argv[1:]]
    dump = None
    if "--so" in args:
        so = args[args.index("--so") + 1]; del args[args.index("--so"):args.index("--so") + 2]
    if "--dump" in args:
        dump = args[args.index("--dump") + 1]; del args[args.index("--dump"):args.index("--dump") + 2]
    if "--build" in args:
        build(so, [a for a in args if a.startswith("-D")]); sys.exit(0)
    sys.argv = [sys.argv[0]] + [a for a in args if not a.startswith("-")]
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
    if dump:
        np.savez(dump, energy=out["energy"], gradient=out["gradient"], qat=out["qat"], niter=out["niter"])
