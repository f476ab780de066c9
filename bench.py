#!/usr/bin/env python
"""Benchmark of the QCxMS production-trajectory hot path on B200 (metric: trajectory-MD-steps/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One "step" = one MD step (leapfrog + one converged cold-start GFN2-xTB energy/gradient + fragment check,
reference src/md.f90:390-454) of EVERY trajectory of the ensemble.  Workload (BASELINE.json configs[1]):
caffeine cation, GFN2-xTB, EI, 1000 trajectories, synthetic initial conditions, exit rules
disabled so every trajectory does identical work.  Trajectories are independent: ranks share nothing during
MD; the only collective is the final histogram all-reduce.

--scaling strong (default; BASELINE config 2 as written): the 1000 trajectories are dealt over the N ranks
(itrj mod N, like bin/pqcxms deals TMP.<n> directories), so the total work is fixed.  --scaling weak: 1000
trajectories PER GPU.  With N > 1 the JSON line carries the other mode's numbers as well ("other_scaling").
At N = 1 the two are the same run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL writes its banner ("NCCL version ...", NCCL_DEBUG=VERSION or WARN) to stdout, which must carry exactly one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "trajectory_md_steps_per_s"
UNIT = "trajectory-MD-steps/s"
E2E_REPS = 3


def algorithmic_flops_per_traj_step(nao, n_it):
    """SURVEY.md 8(d) reporting convention: F_step = n_it (9.7 n^3 + 22 n^2) + 2 n^3 (integrals excluded)."""
    return n_it * (9.7 * nao ** 3 + 22.0 * nao ** 2) + 2.0 * nao ** 3


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)   # gone before the end-to-end leg starts
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_md_sample(num, xyz0, ntraj, nsteps, threads, method=2):
    """The oracle's md() on the host cores: ntraj trajectories x nsteps steps, one trajectory per thread (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as po
    from qcxms_b200 import ensemble_setup as es
    po.lib()
    ic = es.synthetic_initial_conditions(num, xyz0, ntraj)

    def run(k):
        r = po.md(num, ic["mass"], ic["xyz"][k], ic["velo"][k], ic["velof"][k], ic["eimp"][k], ic["tadd"][k], mchrg=1, nmax=nsteps, exit_rules=False, method=method)
        return r["nstep"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        done = sum(ex.map(run, range(ntraj)))
    dt = time.perf_counter() - t0
    return done, dt


def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 4096^3 on this GPU: the FP64 denominator SURVEY.md 8(d) asks for (MEASURED_PEAKS.json has none)."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ntraj", type=int, default=1000, help="trajectories: in total (--scaling strong) or per GPU (--scaling weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--molecule", default="caffeine")
    ap.add_argument("--method", default="gfn2", choices=["gfn2", "gfn1"], help="gfn1: BASELINE config 3 (same molecule, GFN1-xTB Hamiltonian)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--warm-start", action="store_true", help="opt-in fast mode, NOT the reference protocol (SURVEY 8f-4): SCC of a step "
                    "starts from the previous step's converged populations; never the headline number")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from qcxms_b200.api import load_molecule
    num, xyz0, _ = load_molecule(args.molecule)
    method_id = {"gfn2": 2, "gfn1": 1}[args.method]
    nat = len(num)
    cores = len(os.sched_getaffinity(0))
    per_gpu = args.scaling == "weak"
    config = {"workload": "%s cation %s-xTB EI, %d trajectories %s, tstep 0.5 fs, etemp 5000 K, cold-start SCC acc=1.0, exit rules off"
                          % (args.molecule, args.method.upper(), args.ntraj, "per GPU" if per_gpu else "in total, dealt itrj mod N over the ranks"), "nat": nat,
              "ntraj_total": args.ntraj * (args.gpus if per_gpu else 1),
              "l2": "per-step working set (453 KB scratch x resident CTAs + state) is re-streamed every SCC cycle; inputs are not cached between steps"}

    if args.warm_start:
        config["workload"] = config["workload"].replace("cold-start SCC", "WARM-START SCC (opt-in, not the reference protocol)")
        config["warm_start"] = True

    if args.impl == "reference":
        # The reference's own CPU implementation cannot be built here (no Fortran, tblite un-vendored): the oracle port is timed.
        if rank != 0:
            return
        from oracle import pyoracle  # noqa: F401
        ntraj_s, per_step = cores, 8
        for _ in range(max(args.warmup, 1)):
            cpu_md_sample(num, xyz0, min(ntraj_s, 4), 1, cores, method_id)
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            d, _dt = cpu_md_sample(num, xyz0, ntraj_s, per_step, cores, method_id)
            done += d + ntraj_s          # md() starts with one egrad before its first step: same work as a step
        dt = time.perf_counter() - t0
        val = done / dt
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                           "sample": "%d trajectories x %d MD step(s) per bench step (md() incl. its initial egrad), one trajectory per host thread" % (ntraj_s, per_step)},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    import qcxms_b200 as qx
    from qcxms_b200 import ensemble_setup as es
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the path's one collective goes through the library's own NCCL communicator (qcxms_b200_comm_*), set up outside the timed regions
    comm = es.make_comm(local_rank)

    def measure(mode):
        """One full measurement (device-timed K steps + end-to-end leg) of this rank's shard under `mode` scaling."""
        ids = es.shard_indices(args.ntraj, world, rank) if mode == "strong" else np.arange(rank * args.ntraj, (rank + 1) * args.ntraj)
        ntl = len(ids)
        ic = es.synthetic_initial_conditions(num, xyz0, ntl, ids=ids)
        pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in ic.items() if k != "mass"}

        def new_ensemble():
            e = qx.Ensemble(num, ic["mass"], ntl, mchrg=1, tstep_fs=0.5, nmax=10 ** 6, exit_rules=False, device=local_rank, method=method_id)
            if args.warm_start:
                e.set_warm_start(True)
            e.set_all(*[pin[k].numpy() for k in ("xyz", "velo", "velof", "eimp", "tadd")])
            return e

        # ---- kernel-only throughput: inputs resident in HBM, K timed steps after W warm-up steps
        ens = new_ensemble()
        ens.run_md(max_steps=max(args.warmup, 3))
        sampler = ClockSampler(local_rank)
        barrier()
        if rank == 0:
            sampler.start()
        t0 = time.perf_counter()
        steps = ens.run_md(max_steps=args.steps)
        torch.cuda.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if rank == 0 else None
        tim = ens.last_timing()
        dev_s = tim["kernel_ms"] * 1e-3
        tt = torch.tensor([dev_s, wall], dtype=torch.float64, device=dev)
        cnt = torch.tensor([float(steps), float(tim["scc_iterations"]), float(tim["launches"])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dev_s_max, wall_max = tt.tolist()
        total_steps, _scc_unused, launches = cnt.tolist()
        # mean SCC cycles per egrad over the timed steps (scc counter is cumulative over the trajectory's life)
        res0 = ens.results()
        n_it = float(np.mean(res0["scc_iter_total"] / (res0["nstep"] + 1.0)))
        ens.histogram(512)
        ens.close()

        # ---- end to end through the public API with host buffers: H2D of the initial conditions, md(), D2H of the results.
        # Three full repetitions (each one creates its ensemble, uploads, runs, reads back and reduces); the median is reported and
        # all three are kept in the JSON line (box-to-box the first repetition moves by several per cent: allocator and NCCL state)
        e2e_runs, e2e_parts = [], []
        for _rep in range(E2E_REPS):
            barrier()
            t0 = time.perf_counter()
            e2 = new_ensemble()
            t1 = time.perf_counter()
            e2_steps = e2.run_md(max_steps=args.steps)
            t2_ = time.perf_counter()
            e2.results()
            comm.allreduce_histogram(e2, 512)
            torch.cuda.synchronize()
            barrier()
            e2e_wall = time.perf_counter() - t0
            e2e_parts.append({"create_upload_ms": 1e3 * (t1 - t0), "md_ms": 1e3 * (t2_ - t1), "kernel_ms": e2.last_timing()["kernel_ms"],
                              "results_reduce_ms": 1e3 * (t0 + e2e_wall - t2_)})
            e2.close()
            t2 = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
            c2 = torch.tensor([float(e2_steps)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
                dist.all_reduce(c2, op=dist.ReduceOp.SUM)
            e2e_runs.append(c2.item() / t2.item())
        h2d = sum(v.numel() * 8 for v in pin.values())
        d2h = ntl * (nat * (3 * 4 + 1) * 8 + nat * 4 + 14 * 8) + 512 * 8
        return dict(value=total_steps / dev_s_max, dev_s_max=dev_s_max, wall_max=wall_max, total_steps=total_steps, launches=launches, n_it=n_it,
                    clocks=clocks, e2e_value=float(np.median(e2e_runs)), e2e_runs=e2e_runs, e2e_parts=e2e_parts, h2d=h2d, d2h=d2h, ntraj_local=ntl)

    m = measure(args.scaling)
    other = None
    if world > 1:
        om = measure("weak" if args.scaling == "strong" else "strong")
        other = {"scaling": "weak" if args.scaling == "strong" else "strong", "value": om["value"], "e2e": om["e2e_value"],
                 "ms_per_step": 1e3 * om["dev_s_max"] / args.steps, "ntraj_per_gpu": om["ntraj_local"]}
    value, dev_s_max, wall_max, total_steps, launches, n_it, clocks = (m[k] for k in ("value", "dev_s_max", "wall_max", "total_steps", "launches", "n_it", "clocks"))
    e2e_value, h2d, d2h = m["e2e_value"], m["h2d"], m["d2h"]
    config["ntraj_per_gpu"] = m["ntraj_local"]

    comm.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fp64_peak = measure_fp64_peak(torch, dev)
    nsh, nao = {("caffeine", 2): (38, 66), ("caffeine", 1): (48, 76)}.get((args.molecule, method_id), (None, None))
    if nao is None:
        from oracle import pyoracle
        nsh, nao = pyoracle.dims(num, method_id)
    flops = algorithmic_flops_per_traj_step(nao, n_it) * (total_steps / world)   # per GPU
    achieved = flops / dev_s_max / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            per_step = json.load(open(prof)).get("k_md_chunk_dram_bytes_per_traj_step")
            traffic = per_step * (total_steps / world) / max(int(launches / world), 1) if per_step else None
        except Exception:
            traffic = None
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches / world),
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                   "runs": m["e2e_runs"], "parts_rank0": m["e2e_parts"],
                   "note": "median of %d full repetitions of: create + H2D initial conditions (pinned) + md() incl. its initial egrad + D2H of every trajectory's result + histogram all-reduce" % E2E_REPS},
           "roofline": {"bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": traffic,
                        "kernel": "k_md_chunk", "note": "FP64 pipe roofline; peak = cuBLAS DGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry); "
                                                        "achieved = SURVEY 8(d) algorithmic FLOPs, n_it = %.2f SCC cycles/step, nao = %d" % (n_it, nao)},
           "wall_s": wall_max, "scc_cycles_per_step": n_it}
    if other:
        out["other_scaling"] = other
    if not args.no_cpu_baseline and world == 1:
        ncpu_traj, ncpu_steps = cores, 8
        cpu_md_sample(num, xyz0, 2, 1, cores, method_id)
        done, dt = cpu_md_sample(num, xyz0, ncpu_traj, ncpu_steps, cores, method_id)
        out["cpu_baseline"] = {"value": (done + ncpu_traj) / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "%d trajectories x %d MD steps (+1 initial egrad each) of the same workload, oracle md(), one trajectory per host thread" % (ncpu_traj, ncpu_steps)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
