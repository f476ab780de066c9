#!/usr/bin/env python
"""Wall-clock breakdown of the end-to-end leg of bench.py (development aid): constructor, uploads, md(), results, histogram."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qcxms_b200 as qx
from qcxms_b200 import ensemble_setup as es
num, xyz0, _ = qx.load_molecule("caffeine")
ntl, steps = 1000, 20
ic = es.synthetic_initial_conditions(num, xyz0, ntl, ids=np.arange(ntl))
comm = es.make_comm(0)
for rep in range(3):
    t = [time.perf_counter()]
    e = qx.Ensemble(num, ic["mass"], ntl, mchrg=1, tstep_fs=0.5, nmax=10 ** 6, exit_rules=False, device=0, method=2); t.append(time.perf_counter())
    e.set_all(*[ic[k] for k in ("xyz", "velo", "velof", "eimp", "tadd")]); t.append(time.perf_counter())
    e.run_md(max_steps=steps); t.append(time.perf_counter())
    e.results(); t.append(time.perf_counter())
    comm.allreduce_histogram(e, 512); torch.cuda.synchronize(); t.append(time.perf_counter())
    tim = e.last_timing()
    t[-1] = time.perf_counter(); e.close(); t.append(time.perf_counter())
    print("rep %d: ctor %.1f  set_all %.1f  run_md %.1f (kernel %.1f)  results %.1f  histogram %.1f  close %.1f ms" %
          ((rep,) + tuple(1e3 * (t[i + 1] - t[i]) for i in range(3)) + (tim.get("kernel_ms", float("nan")),) + tuple(1e3 * (t[i + 1] - t[i]) for i in range(3, 6))), flush=True)
