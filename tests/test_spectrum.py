"""Spectrum assembly from qcxms.res records: record parser, isotope patterns, JCAMP reader, similarity."""
import os

import numpy as np

from qcxms_b200 import fragments as fr
from qcxms_b200 import spectrum as sp

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_parse_record_round_trips_res_line():
    pairs = [(1, 5), (6, 2), (8, 1), (17, 1)]
    chg, mchrg, got = sp.parse_record(fr.res_line(0.7312345, 1, 17, 2, 1, pairs))
    assert abs(chg - 0.7312345) < 1e-7 and mchrg == 1 and got == pairs
    chg, mchrg, got = sp.parse_record(fr.res_line(1.0, 1, 17, 2, 1, pairs, icoll=3))
    assert chg == 1.0 and got == pairs
    assert sp.parse_record(fr.res_line(0.5, 1, 3, 1, 2, [(113, 1), (1, 3)]))[2] == [(113, 1), (1, 3)]


def test_isotope_pattern_known_cases():
    # one chlorine: 35 / 37 in the natural 3:1 ratio
    p = sp.isotope_pattern([(17, 1)], 64)
    assert abs(p[35] - 0.7576) < 1e-12 and abs(p[37] - 0.2424) < 1e-12 and abs(p.sum() - 1.0) < 1e-12
    # two chlorines: 9 : 6 : 1 (binomial)
    p = sp.isotope_pattern([(17, 2)], 128)
    assert np.allclose([p[70], p[72], p[74]], [0.7576 ** 2, 2 * 0.7576 * 0.2424, 0.2424 ** 2], atol=1e-14)
    # caffeine C8H10N4O2: M = 194, M+1 / M about 10 % (8 x 1.08 % 13C + 4 x 0.37 % 15N + ...)
    p = sp.isotope_pattern([(1, 10), (6, 8), (7, 4), (8, 2)], 256)
    assert np.argmax(p) == 194 and abs(p.sum() - 1.0) < 1e-12
    assert 0.095 < p[195] / p[194] < 0.11
    # an explicit isotope label (100 + mass) is a single line
    p = sp.isotope_pattern([(113, 1), (1, 4)], 64)
    assert np.argmax(p) == 17 and abs(p[17] - 0.999885 ** 4) < 1e-12
    # binary-power convolution == repeated convolution
    q = np.zeros(256); q[0] = 1.0
    e = np.zeros(256); e[12], e[13] = 0.9893, 0.0107
    for _ in range(11):
        q = np.convolve(q, e)[:256]
    assert np.allclose(sp.isotope_pattern([(6, 11)], 256), q, atol=1e-15)


def test_spectrum_is_charge_weighted_and_additive():
    a = [fr.res_line(0.75, 1, 1, 1, 1, [(1, 3), (6, 1)]), fr.res_line(0.25, 1, 1, 1, 2, [(1, 2), (6, 1), (8, 1), (17, 1)])]
    b = [fr.res_line(1.0, 1, 2, 1, 1, [(1, 5), (6, 2), (8, 1), (17, 1)])]
    sa, sb, sab = sp.spectrum(a, 128), sp.spectrum(b, 128), sp.spectrum(a + b, 128)
    assert np.allclose(sa + sb, sab, atol=1e-15)              # ranks can add their arrays (the all-reduce)
    assert abs(sab.sum() - 2.0) < 1e-9                        # total statistical charge is conserved
    assert abs(sa[15] - 0.75 * 0.9893 * 0.999885 ** 3) < 1e-12
    stick = sp.spectrum(a + b, 128, isotopes=False)
    assert stick[15] == 0.75 and stick[65] == 0.25 and stick[80] == 1.0
    assert np.allclose(stick, fr.spectrum_from_records(a + b, 128))
    assert sab[82] / sab[80] > 0.3                            # the 37Cl satellite of the molecular ion


def test_reference_experimental_spectra_read_and_compare():
    e = sp.read_jcamp(os.path.join(GOLD, "exp_2-chloroethanol.jdx"))
    assert e[31] == 9999 and e[80] == 390 and e[82] == 120 and np.count_nonzero(e) == 48
    m = sp.read_jcamp(os.path.join(GOLD, "exp_monoethanolamine.jdx"))
    assert np.argmax(m) == 30
    assert abs(sp.cosine_similarity(e, e) - 1.0) < 1e-15
    assert sp.cosine_similarity(e, m) < 0.5
    # a two-line model of the chloroethanol spectrum (CH2OH+ base peak, molecular ion) already correlates with the experiment
    model = sp.spectrum([fr.res_line(0.9, 1, 1, 1, 1, [(1, 3), (6, 1), (8, 1)]), fr.res_line(0.1, 1, 2, 1, 1, [(1, 5), (6, 2), (8, 1), (17, 1)])])
    assert sp.cosine_similarity(sp.normalise(model), e) > 0.8
