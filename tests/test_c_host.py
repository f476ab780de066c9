"""The C ABI driven from a plain C program the way the Fortran shim would drive it, on a reference-format trajectory directory
(start.xyz + qcxms.start, src/utility.f90:363-422)."""
import os
import subprocess

import numpy as np
import pytest

from qcxms_b200 import startfiles as sf
from qcxms_b200.api import load_molecule

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _build_host(tmp_path):
    exe = str(tmp_path / "host_ei")
    libdir = os.path.join(ROOT, "qcxms_b200")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "c_host", "host_ei.c"),
                           "-L", libdir, "-lqcxms_b200", "-Wl,-rpath," + os.path.abspath(libdir), "-lm"])
    return exe


def test_fortran_d_format_and_round_trip(tmp_path):
    assert sf.fortran_d(1.0) == "  0.10000000000000D+01" and sf.fortran_d(-0.012345678901234) == " -0.12345678901234D-01"
    assert sf.fortran_d(0.0) == "  0.00000000000000D+00" and sf.fortran_d(0.99999999999999999) == "  0.10000000000000D+01"
    num, xyz, _ = load_molecule("chloroethanol")
    rng = np.random.default_rng(1)
    velo, velof = 1e-3 * rng.standard_normal((len(num), 3)), np.ones(len(num))
    sf.write_start(str(tmp_path), 17, num, xyz, velo, velof, 0.31234, 16536.5)
    lines = open(tmp_path / "qcxms.start").read().splitlines()
    assert lines[0] == "  17" and len(lines) == 3 + len(num) and all(len(ln) == 88 for ln in lines[3:])
    back = sf.read_start(str(tmp_path))
    assert back["itrj"] == 17 and np.array_equal(back["num"], num)
    assert np.abs(back["xyz"] - xyz).max() < 1e-12 and np.abs(back["velo"] - velo).max() < 1e-16
    assert abs(back["eimp"] - 0.31234) < 1e-15 and abs(back["tadd"] - 16536.5) < 1e-9


def test_c_host_compiles_and_links_against_the_abi(tmp_path):
    """gcc sees exactly the declarations of include/qcxms_b200.h; the program must link against the built library and fail cleanly
    (no GPU here) instead of crashing."""
    exe = _build_host(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "nowhere"), "5"], capture_output=True, text=True)
    assert r.returncode == 1 and "Missing start.xyz" in r.stderr


@pytest.mark.gpu
def test_c_host_runs_a_reference_format_directory(qx, oracle, tmp_path):
    from qcxms_b200 import ensemble_setup as es
    num, xyz, _ = qx.load_molecule("chloroethanol")
    ic = es.synthetic_initial_conditions(num, xyz, 1)
    sf.write_start(str(tmp_path), 1, num, ic["xyz"][0], ic["velo"][0], ic["velof"][0], ic["eimp"][0], ic["tadd"][0])
    exe = _build_host(tmp_path)
    out = subprocess.check_output([exe, str(tmp_path), "15"], text=True).splitlines()
    head = out[0].split()
    assert int(head[head.index("nstep") + 1]) == 15 and int(head[head.index("mdok") + 1]) == 1
    epot = float(out[1].split()[1])
    got_xyz = np.array([[float(v) for v in ln.split()[2:5]] for ln in out[2:2 + len(num)]])
    # the same trajectory through the oracle, from what the files hold (14 digits)
    st = sf.read_start(str(tmp_path))
    ref = oracle.md(num, ic["mass"], st["xyz"], st["velo"], st["velof"], st["eimp"], st["tadd"], mchrg=1, nmax=15)
    assert abs(epot - ref["Epot"]) < 1e-7 and np.abs(got_xyz - ref["xyz"]).max() < 1e-6
