"""Host logic after md(): fragment IPs, statistical charge assignment, qcxms.res records."""
import numpy as np

from qcxms_b200 import ensemble_setup as es
from qcxms_b200 import fragments as fr
from qcxms_b200.api import load_molecule


def test_res_line_fortran_format(oracle):
    # '(F10.7,i3,2i5,2i2,2x,i3,2x,20(i4,i3))' -- EI passes one integer fewer than CID, so its columns are shifted
    ei = fr.res_line(0.1234567, 1, 12, 1, 2, [(1, 3), (6, 2)])
    assert ei == " 0.1234567  1   12    1 2 2    1     3  6   2"
    cid = fr.res_line(1.0, 1, 12, 1, 2, [(1, 3), (6, 2)], icoll=3)
    assert cid == " 1.0000000  1   12    3 1 2    2     1  3   6  2"
    assert fr.res_line(-0.5, -1, 7, 2, 1, [(117, 1)]) == "-0.5000000 -1    7    2 1 1  117     1"
    rng = np.random.default_rng(4)
    for _ in range(200):
        pairs = sorted({int(z): int(c) for z, c in zip(rng.integers(1, 136, 6), rng.integers(1, 30, 6))}.items())
        args = (float(rng.random()), 1, int(rng.integers(1, 120000)), int(rng.integers(1, 8)), int(rng.integers(1, 6)), pairs)
        icoll = None if rng.random() < 0.5 else int(rng.integers(1, 12))
        assert fr.res_line(*args, icoll=icoll) == oracle.res_line(*args, icoll=icoll)


def test_boltz_matches_oracle(oracle):
    ip = np.array([9.8, 10.4, 13.6])
    p = fr.boltz(2, 4000.0, ip)
    assert abs(p.sum() - 1.0) < 1e-15 and p[0] > p[1] > p[2]
    assert np.allclose(p, oracle.boltz(4000.0, ip), rtol=1e-14, atol=0)
    kt_ev = 4000.0 * 3.166808578545117e-06 * 27.21138505
    assert abs(p[1] / p[0] - np.exp(-(10.4 - 9.8) / kt_ev)) < 1e-14


def test_eself_and_electron_counting():
    assert fr.electrons_amount([1], 1)[0] == 0 and fr.electrons_amount([17, 1], 0)[0] == 8
    assert abs(fr.eself([1], [1.0]) - (0.5 * 0.47259288 - 0.02448 / 3.0)) < 1e-15
    assert fr.getspin([1], 1) == -1 and fr.getspin([6, 1, 1, 1], 0) == 2 and fr.getspin([8, 1, 1], 0) == 1


def test_avg_frag_struc_and_pairs():
    num = np.array([6, 1, 17, 1, 8]); lst = np.array([1, 1, 2, 1, 1])
    xyz = np.arange(15.0).reshape(5, 3)
    natf, iatf, xyzf = fr.avg_frag_struc(num, xyz, lst, 2)
    assert natf == [4, 1] and iatf[0].tolist() == [6, 1, 1, 8] and iatf[1].tolist() == [17]
    assert np.array_equal(xyzf[1][0], xyz[2])
    assert fr.fragat_pairs(num, lst, 1) == [(1, 2), (6, 1), (8, 1)]
    assert fr.fragat_pairs(num, lst, 1, imass=[13, 0, 0, 2, 0]) == [(1, 1), (8, 1), (102, 1), (113, 1)]


def test_manage_fragments_c_cl_cleavage(oracle):
    """2-chloroethanol cation with the chlorine pulled off: two fragments, IPs from the (CPU) single points, the charge
    goes (almost entirely) to the fragment with the lower IP and the other record is written at once."""
    num, xyz, _ = load_molecule("chloroethanol")
    mass = es.masses_au(num)
    x = xyz.copy(); x[2] += np.array([9.0, -6.0, -6.0])
    lst = oracle.fragment_structure(num, x, 3.0)
    assert lst.max() == 2
    out = fr.manage_fragments(num, mass, x, lst, np.full(len(num), 1.0 / len(num)), aTlast=3000.0, itrj=5, isec=1, energies=oracle.energies)
    assert out["nfrag"] == 2 and out["nfrag_ok"] and out["ipok"]
    ip = out["fragip"]
    assert 5.0 < ip.min() < ip.max() < 20.0                       # eV, plausible IPs of C2H5O / Cl
    assert abs(out["fragchrg3"].sum() - 1.0) < 1e-12
    assert np.allclose(out["fragchrg3"], oracle.boltz(3000.0, ip))
    assert out["tcont"] == 1 + int(np.argmax(np.array(out["natf"]) * out["fragchrg3"]))
    assert len(out["lines"]) == 1 and out["asave"] is not None and out["mchrg"] == 1
    other = 2 if out["tcont"] == 1 else 1
    assert out["lines"][0] == oracle.res_line(out["fragchrg3"][other - 1], 1, 5, 1, other, fr.fragat_pairs(num, lst, other))
    assert abs(out["chrgcont"] - out["fragchrg3"][out["tcont"] - 1]) < 1e-15
    # masses as fragmass reports them (amu)
    assert abs(sum(out["fragm"]) - 80.5) < 0.2
    # one fragment only: nothing is held back, the full charge stays
    one = fr.manage_fragments(num, mass, xyz, np.ones(len(num), dtype=np.int32), np.zeros(len(num)), aTlast=500.0, itrj=5, isec=1, energies=oracle.energies)
    assert one["tcont"] == 0 and one["asave"] is None and len(one["lines"]) == 1 and one["lines"][0].startswith(" 1.0000000  1    5    1 1 4")


def test_spectrum_from_records():
    lines = [fr.res_line(0.25, 1, 1, 1, 2, [(17, 1)]), fr.res_line(0.75, 1, 1, 1, 1, [(1, 5), (6, 2), (8, 1)])]
    bins = fr.spectrum_from_records(lines, 128)
    assert bins[35] == 0.25 and bins[45] == 0.75 and bins.sum() == 1.0
