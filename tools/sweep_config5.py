#!/usr/bin/env python
"""BASELINE config 5: ensemble-size sweep (256 ... 8192 trajectories) of the ~100-atom molecule through bench.py.

    python tools/sweep_config5.py [--gpus N] [--sizes 256,512,...] [--steps K]   -> one bench JSON line per size (gpurun_out/)
"""
import argparse, json, os, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--sizes", default="256,512,1024,2048,4096,8192")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--molecule", default="peptide_cl")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config5_sweep.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "a") as f:
        for n in [int(v) for v in a.sizes.split(",")]:
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", str(a.gpus), "--steps", str(a.steps), "--warmup", "3", "--molecule", a.molecule,
                   "--ntraj", str(n), "--no-cpu-baseline"]
            if a.gpus > 1:
                cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus), "--master-addr", "127.0.0.1",
                       "--master-port", "29533"] + cmd[1:]
            r = subprocess.run(cmd, capture_output=True, text=True)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if not line:
                print("size", n, "FAILED", r.stderr[-400:], flush=True)
                continue
            d = json.loads(line[-1])
            print("ntraj %5d gpus %d: %.1f steps/s (e2e %.1f), %.1f ms/step, frac %.4f" % (n, a.gpus, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]), flush=True)
            f.write(line[-1] + "\n"); f.flush()


if __name__ == "__main__":
    main()
