#!/usr/bin/env python
"""Small cases of every kernel family for compute-sanitizer (no torch import):
    compute-sanitizer --tool memcheck python tools/sanitize_cases.py [cases...]
cases: small wide medium large md mdwide gfn1 (default: all)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qcxms_b200 as qx
from qcxms_b200 import ensemble_setup as es

cases = sys.argv[1:] or ["small", "wide", "medium", "large", "md", "mdwide", "gfn1"]
rng = np.random.default_rng(0)
NBIG = int(os.environ.get("QX_SAN_N", "300"))   # systems of the "more systems than CTA slots" cases (racecheck: use a few)


def single_points(name, nsys, method=qx.gfn2_xtb):
    num, xyz, _ = qx.load_molecule(name)
    geoms = xyz[None] + 0.03 * rng.standard_normal((nsys,) + xyz.shape)
    out = qx.egrad_batch(num, geoms, 1, 2, method, 5000.0)
    assert (out["stat"] == 0).all(), out["stat"]
    return float(out["energy"][0])


def md(name, ntraj, steps):
    num, xyz, _ = qx.load_molecule(name)
    ic = es.synthetic_initial_conditions(num, xyz, ntraj, first_id=0)
    ens = qx.Ensemble(num, ic["mass"], ntraj, mchrg=1, nmax=steps, exit_rules=True)
    ens.set_all(ic["xyz"], ic["velo"], ic["velof"], ic["eimp"], ic["tadd"])
    n = ens.run_md()
    r = ens.results()
    ens.histogram(64)
    ens.close()
    return n, int((r["status"] == 2).sum())


for c in cases:
    t0 = time.perf_counter()
    if c == "small":      # two 320-thread CTAs per SM: more systems than SMs
        got = single_points("chloroethanol", NBIG)
    elif c == "wide":     # 576-thread CTAs + eigenpair refinement (no more systems than SMs)
        got = single_points("caffeine", 3)
    elif c == "medium":   # 512-thread CTAs, shared-memory matrices, multi-pass Jacobi
        got = single_points("alkane_c14", 2)
    elif c == "large":    # 512-thread CTAs, matrices in the global slab, blocked Jacobi over the overlaid block buffer
        got = single_points("peptide_cl", 1)
    elif c == "gfn1":
        got = single_points("monoethanolamine", 3, qx.gfn1_xtb)
    elif c == "md":       # persistent MD kernel, narrow CTAs
        got = md("chloroethanol", NBIG, 6)
    elif c == "mdwide":   # persistent MD kernel, wide CTAs with eigenvector seeds
        got = md("caffeine", 3, 6)
    else:
        raise SystemExit("unknown case " + c)
    print("case %-7s ok: %s  (%.1f s)" % (c, got, time.perf_counter() - t0), flush=True)
