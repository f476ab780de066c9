"""Mass spectrum from qcxms.res / qcxms_cid.res records.

The reference writes one record per charged fragment -- statistical charge, charge state, trajectory, [collision,] cascade
level, fragment index and the element (or 100 + isotope mass) counts (format of src/write_fragments.f90:402-441) -- and leaves
the spectrum to PlotMS, which is not part of the reference tree (README.md:77-90): it sums the statistical charges per
sum formula, expands every formula into its isotope pattern and normalises the largest peak.  This module does that on
the host:

* `parse_record`      one record -> (charge, mchrg, [(Z or 100 + isotope mass, count)])
* `isotope_pattern`   exact pattern of a formula by polynomial convolution of the natural abundances (nominal masses)
* `spectrum`          charge-weighted, isotope-expanded intensities per integer m/z -- the quantity the ranks all-reduce
* `read_jcamp`        the experimental spectra the reference ships (share/examples/EI/*/exp.dat, JCAMP-DX peak tables)
* `cosine_similarity` the comparison BASELINE.json's north star names for whole-ensemble spectra

PlotMS draws the isotope pattern by random sampling; the exact convolution used here is its infinite-sample limit.
"""
import numpy as np

# natural isotopic compositions (nominal mass, abundance), H .. Ar and the halogens the reference's examples use
ISOTOPES = {
    1: [(1, 0.999885), (2, 0.000115)],
    2: [(3, 0.00000134), (4, 0.99999866)],
    3: [(6, 0.0759), (7, 0.9241)],
    4: [(9, 1.0)],
    5: [(10, 0.199), (11, 0.801)],
    6: [(12, 0.9893), (13, 0.0107)],
    7: [(14, 0.99636), (15, 0.00364)],
    8: [(16, 0.99757), (17, 0.00038), (18, 0.00205)],
    9: [(19, 1.0)],
    10: [(20, 0.9048), (21, 0.0027), (22, 0.0925)],
    11: [(23, 1.0)],
    12: [(24, 0.7899), (25, 0.1000), (26, 0.1101)],
    13: [(27, 1.0)],
    14: [(28, 0.92223), (29, 0.04685), (30, 0.03092)],
    15: [(31, 1.0)],
    16: [(32, 0.9499), (33, 0.0075), (34, 0.0425), (36, 0.0001)],
    17: [(35, 0.7576), (37, 0.2424)],
    18: [(36, 0.003365), (38, 0.000632), (40, 0.996003)],
    35: [(79, 0.5069), (81, 0.4931)],
    53: [(127, 1.0)],
}


def parse_record(line):
    """(charge, mchrg, pairs) of one qcxms.res (EI) or qcxms_cid.res (CID: one more integer, icoll) record."""
    chg = float(line[:10])
    vals = [int(v) for v in line[10:].split()]
    for pos in (4, 5):   # mchrg itrj [icoll] isec ifrag ntypes (Z count)*ntypes
        if pos < len(vals) and len(vals) - pos - 1 == 2 * vals[pos]:
            n = vals[pos]
            return chg, vals[0], [(vals[pos + 1 + 2 * k], vals[pos + 2 + 2 * k]) for k in range(n)]
    raise ValueError("not a qcxms.res record: %r" % line)


def isotope_pattern(pairs, nbins):
    """Probability of every integer mass of the formula [(Z or 100 + isotope mass, count), ...] (sums to 1 unless it leaves nbins)."""
    p = np.zeros(nbins)
    p[0] = 1.0
    for z, count in pairs:
        iso = [(z - 100, 1.0)] if z > 100 else ISOTOPES[z]
        elem = np.zeros(nbins)
        for m, a in iso:
            if m < nbins:
                elem[m] = a
        # count-fold convolution by binary powers
        acc, base, c = None, elem, int(count)
        while c:
            if c & 1:
                acc = base.copy() if acc is None else np.convolve(acc, base)[:nbins]
            c >>= 1
            if c:
                base = np.convolve(base, base)[:nbins]
        if acc is not None:
            p = np.convolve(p, acc)[:nbins]
    return p


def spectrum(lines, nbins=512, isotopes=True):
    """Intensity per integer m/z: sum over the records of |statistical charge| x isotope pattern, at m / |mchrg|.
    Not normalised (ranks add their arrays; `normalise` afterwards)."""
    out = np.zeros(nbins)
    cache = {}
    for ln in lines:
        if not ln.strip():
            continue
        chg, mchrg, pairs = parse_record(ln)
        key = tuple(pairs)
        pat = cache.get(key)
        if pat is None:
            if isotopes:
                pat = isotope_pattern(pairs, nbins)
            else:
                pat = np.zeros(nbins)
                m = sum((z - 100 if z > 100 else max(ISOTOPES[z], key=lambda t: t[1])[0]) * c for z, c in pairs)
                if m < nbins:
                    pat[m] = 1.0
            cache[key] = pat
        z = max(abs(int(mchrg)), 1)
        if z == 1:
            out += abs(chg) * pat
        else:
            for m in np.nonzero(pat)[0]:
                out[int(round(m / z))] += abs(chg) * pat[m]
    return out


def normalise(bins, top=1000.0):
    """Largest peak = top (PlotMS writes 0..1000 into its JCAMP file; NIST tables use 9999)."""
    mx = bins.max()
    return bins * (top / mx) if mx > 0 else bins.copy()


def read_jcamp(path, nbins=512):
    """Peak table of a JCAMP-DX mass spectrum (##PEAK TABLE=(XY..XY) ... ##END=) as an intensity array over integer m/z."""
    bins = np.zeros(nbins)
    on = False
    for ln in open(path):
        s = ln.strip()
        if s.startswith("##"):
            on = s.upper().startswith("##PEAK TABLE") or s.upper().startswith("##XYDATA") or s.upper().startswith("##XYPOINTS")
            continue
        if not on or not s:
            continue
        for tok in s.replace(";", " ").split():
            xy = tok.split(",")
            if len(xy) == 2:
                m = int(round(float(xy[0])))
                if 0 <= m < nbins:
                    bins[m] += float(xy[1])
    return bins


def cosine_similarity(a, b, mz_min=0):
    a = np.asarray(a, dtype=np.float64)[mz_min:]
    b = np.asarray(b, dtype=np.float64)[mz_min:]
    n = min(len(a), len(b))
    den = np.linalg.norm(a[:n]) * np.linalg.norm(b[:n])
    return float(a[:n] @ b[:n] / den) if den > 0 else 0.0
