"""GFN1-xTB restatement (method id 1, reference src/tblite.f90:124-126; BASELINE config 3): self-consistency of the oracle.
The element parameters are UNVERIFIED (params/gfn1_params.h) -- these tests pin the structure of the method, not its digits."""
import os
import subprocess
import sys

import numpy as np
import pytest

from qcxms_b200.api import load_molecule

AA = 1.0 / 0.52917726
# CH3Cl ... NH3 with a near-linear C-Cl...N contact: exercises the halogen-bond correction
XB_NUM = np.array([6, 17, 1, 1, 1, 7, 1, 1, 1], dtype=np.int32)
XB_XYZ = np.array([[0, 0, 0], [1.78, 0, 0], [-0.36, 1.03, 0], [-0.36, -0.51, 0.89], [-0.36, -0.51, -0.89], [4.9, 0.1, 0.05],
                   [5.27, 0.95, 0.1], [5.27, -0.40, 0.85], [5.27, -0.35, -0.80]]) * AA


def test_dimensions_follow_the_survey(oracle):
    # H carries two s shells in GFN1 => monoethanolamine nao = 30, caffeine nao = 76
    assert oracle.dims(load_molecule("monoethanolamine")[0], 1) == (22, 30)
    assert oracle.dims(load_molecule("caffeine")[0], 1) == (48, 76)
    assert oracle.dims(load_molecule("caffeine")[0], 2) == (38, 66)
    with pytest.raises(ValueError):
        oracle.dims(np.array([18, 1], dtype=np.int32), 1)      # no built-in GFN1 parameters for Ar


@pytest.mark.parametrize("case", ["monoethanolamine", "chloroethanol", "xb"])
def test_gradient_matches_finite_differences(oracle, case):
    if case == "xb":
        num, x, charge, mult, etemp = XB_NUM, XB_XYZ, 0, 1, 300.0
    else:
        num, xyz, _ = load_molecule(case)
        x = xyz + 0.06 * np.random.default_rng(3).standard_normal(xyz.shape)
        charge, mult, etemp = 1, 2, 5000.0
    oracle.set_accuracy(1e-4)
    try:
        r = oracle.egrad(num, x, charge, mult, 1, etemp)
        assert r["stat"] == 0
        g, h = r["gradient"], 1e-4
        fd = np.zeros_like(x)
        for i in range(len(num)):
            for c in range(3):
                xp, xm = x.copy(), x.copy()
                xp[i, c] += h; xm[i, c] -= h
                fd[i, c] = (oracle.egrad(num, xp, charge, mult, 1, etemp)["energy"] - oracle.egrad(num, xm, charge, mult, 1, etemp)["energy"]) / (2 * h)
    finally:
        oracle.set_accuracy(1.0)
    assert np.abs(g - fd).max() < 2e-7
    assert np.abs(g.sum(0)).max() < 1e-10


def test_invariances_and_charge_conservation(oracle):
    num, xyz, _ = load_molecule("monoethanolamine")
    oracle.set_accuracy(1e-3)
    try:
        a = oracle.egrad(num, xyz, 1, 2, 1, 5000.0)
        th = 0.7
        R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]]) @ np.array([[1, 0, 0], [0, np.cos(0.3), -np.sin(0.3)], [0, np.sin(0.3), np.cos(0.3)]])
        b = oracle.egrad(num, xyz @ R.T + np.array([3.0, -2.0, 0.5]), 1, 2, 1, 5000.0)
        perm = np.random.default_rng(1).permutation(len(num))
        c = oracle.egrad(num[perm], xyz[perm], 1, 2, 1, 5000.0)
    finally:
        oracle.set_accuracy(1.0)
    assert abs(a["energy"] - b["energy"]) < 1e-10 and abs(a["energy"] - c["energy"]) < 1e-10
    assert np.abs(a["gradient"] @ R.T - b["gradient"]).max() < 1e-9
    assert np.abs(a["qat"][perm] - c["qat"]).max() < 1e-9
    assert abs(a["qat"].sum() - 1.0) < 1e-8


def test_model_pieces(oracle):
    num, xyz, _ = load_molecule("monoethanolamine")
    r = oracle.egrad(num, xyz, 0, 1, 1, 300.0, detail=True)
    assert r["stat"] == 0 and r["converged"] == 1
    assert r["e_aes"] == 0.0 and r["e_disp_sc"] == 0.0            # no multipoles, no self-consistent dispersion
    assert -0.01 < r["e_disp_atm"] < -0.001                          # D3(BJ), in-tree reference data
    S = r["overlap"]
    assert np.abs(np.diag(S) - 1.0).max() < 1e-12                    # every function normalised, also the orthogonalised H 2s
    ih = int(np.where(num == 1)[0][0])
    ao = np.where(r["ao2at"] == ih)[0]
    assert len(ao) == 2 and abs(S[ao[0], ao[1]]) < 1e-12            # H 2s is orthogonal to H 1s
    # D3 coordination numbers: about 4 for the carbons, about 1 for the hydrogens
    assert np.all(np.abs(r["cn"][num == 6] - 4.0) < 0.3) and np.all(np.abs(r["cn"][num == 1] - 1.0) < 0.15)
    # the example geometry is (close to) a minimum of the method: a wrong term or element row would leave forces of 1e-2 .. 1e-1
    assert np.abs(r["gradient"]).max() < 8e-3
    # halogen bond: attractive, and gone without an acceptor in range
    e_with = oracle.egrad(XB_NUM, XB_XYZ, 0, 1, 1, 300.0)["energy"]
    far = XB_XYZ.copy(); far[5:] += np.array([40.0, 0, 0]) * AA
    e_far = oracle.egrad(XB_NUM, far, 0, 1, 1, 300.0)["energy"]
    assert e_with < e_far


def test_parameter_file_override(tmp_path):
    """QCXMS_B200_GFN1_PARAM: an xtb-format parameter file replaces the built-in table (here: the built-in H and O rows written
    out again plus a shifted O level, in a fresh process)."""
    f = tmp_path / "param_gfn1.txt"
    f.write_text("$globpar\n ks 1.85\n$end\n$Z= 1\n ao=1s2s\n lev= -10.923452 -2.171902\n exp= 1.207940 1.993207\n EN=2.20\n GAM=0.470099\n GAM3=0.0\n"
                 " REPA=2.209700\n REPB=1.116244\n POLYS=0.0\n$end\n$Z= 8\n ao=2s2p\n lev= -23.398376 -17.886554\n exp= 2.345365 2.153060\n EN=3.44\n"
                 " GAM=0.583349\n GAM3=-0.517134\n REPA=2.004253\n REPB=5.171786\n POLYS=-13.729047\n POLYP=-4.453341\n LPARP=0.451896\n$end\n")
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle import pyoracle as po;"
            "num = np.array([8, 1, 1], dtype=np.int32); xyz = np.array([[0, 0, 0], [1.8, 0, 0], [-0.45, 1.75, 0.0]]);"
            "print(repr(po.egrad(num, xyz, 0, 1, 1, 300.0)['energy']))" % os.path.join(os.path.dirname(__file__), ".."))
    env = dict(os.environ)
    env.pop("QCXMS_B200_GFN1_PARAM", None)
    e0 = float(subprocess.check_output([sys.executable, "-c", code], env=env).decode().split()[-1])
    env["QCXMS_B200_GFN1_PARAM"] = str(f)
    e1 = float(subprocess.check_output([sys.executable, "-c", code], env=env).decode().split()[-1])
    assert abs(e0 - e1) < 1e-12                                      # the same numbers through the file
    f.write_text(f.read_text().replace("-23.398376", "-23.898376"))
    e2 = float(subprocess.check_output([sys.executable, "-c", code], env=env).decode().split()[-1])
    assert abs(e2 - e0) > 1e-3
