"""Host-side mirror of the reference interface of the hot path, on top of the C ABI.

Names follow the reference: get_xtb_egrad (src/tblite.f90:65), the method selectors gfn1_xtb /
gfn2_xtb / ipea1_xtb (src/tblite.f90:34-40), md() results (src/md.f90:34-39), fragment_structure
(src/fragments.f90:93).  All heavy lifting happens in libqcxms_b200.so (hand-written sm_100a CUDA);
this module only marshals numpy arrays through ctypes.
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QCXMS_B200_LIB") or os.path.join(_HERE, "libqcxms_b200.so")   # (the override is a measurement hook)
_LIB = None

# method selectors (reference src/tblite.f90:34-40)
gfn1_xtb, gfn2_xtb, ipea1_xtb = 1, 2, 11

FSTOAU = 41.3413733365614       # reference src/xtb_mctc_convert.f90:55
AUTOEV = 27.21138505
AMUTOAU = 1.660539040e-27 * (1.0 / 9.10938356e-31)
KB = 3.166808578545117e-06


class MdConfig(C.Structure):
    """qcxms_b200_md_config_t"""
    _fields_ = [("method_id", C.c_int32), ("mchrg", C.c_int32), ("nfragexit", C.c_int32), ("exit_rules", C.c_int32),
                ("nmax", C.c_int32), ("isec", C.c_int32), ("tstep", C.c_double), ("etemp_in", C.c_double),
                ("ieetemp", C.c_double), ("ax", C.c_double)]


class MdResult(C.Structure):
    """qcxms_b200_md_result_t"""
    _fields_ = [("mdok", C.c_int32), ("fragstate", C.c_int32), ("nstep", C.c_int32), ("nfrag", C.c_int32),
                ("status", C.c_int32), ("scc_iter_total", C.c_int32)] + \
        [(k, C.c_double) for k in ("Tav", "Epav", "Ekav", "aTlast", "dtime", "ttime", "Epot", "Ekin")]


class CidConfig(C.Structure):
    """qcxms_b200_cid_config_t"""
    _fields_ = [("method_id", C.c_int32), ("mchrg", C.c_int32), ("gas_z", C.c_int32), ("eexact", C.c_int32),
                ("manual_dist", C.c_int32), ("ntot", C.c_int32), ("gas_mass", C.c_double), ("tstep", C.c_double),
                ("etemp", C.c_double), ("elab", C.c_double), ("ecom", C.c_double)]


class CidResult(C.Structure):
    """qcxms_b200_cid_result_t"""
    _fields_ = [("stopcid", C.c_int32), ("nstep", C.c_int32), ("nfrag", C.c_int32), ("collided", C.c_int32),
                ("status", C.c_int32), ("scc_iter_total", C.c_int32), ("velo_cm", C.c_double), ("aTlast", C.c_double),
                ("ttime", C.c_double), ("epot", C.c_double), ("direc", C.c_double * 3)]


# collision gases of the reference (src/input.f90:512-558): Z, mass / amu, radius / bohr
GASES = {"he": (2, 4.002, 2.64560263), "ne": (10, 20.18, 2.91016289), "ar": (18, 39.948, 3.55266638), "n2": (7, 14.007, 3.64)}

EXPORTS = ["qcxms_b200_egrad", "qcxms_b200_egrad_spec", "qcxms_b200_basis_size", "qcxms_b200_cid_batch", "qcxms_b200_egrad_batch", "qcxms_b200_fragment_structure", "qcxms_b200_ensemble_create",
           "qcxms_b200_ensemble_destroy", "qcxms_b200_ensemble_set_trajectory", "qcxms_b200_ensemble_set_all",
           "qcxms_b200_ensemble_run_md", "qcxms_b200_ensemble_set_warm_start", "qcxms_b200_ensemble_set_mfp", "qcxms_b200_ensemble_get_new_velo", "qcxms_b200_ensemble_get_result", "qcxms_b200_ensemble_get_all", "qcxms_b200_ensemble_last_timing",
           "qcxms_b200_ensemble_histogram", "qcxms_b200_ensemble_intenergy", "qcxms_b200_last_error", "qcxms_b200_version",
           "qcxms_b200_comm_unique_id", "qcxms_b200_comm_create", "qcxms_b200_comm_destroy", "qcxms_b200_comm_allreduce_sum",
           "qcxms_b200_ensemble_allreduce_histogram", "qcxms_b200_ensemble_set_gs_mode", "qcxms_b200_ensemble_get_gs", "qcxms_b200_ensemble_set_esi"]


def lib():
    """Load libqcxms_b200.so; fails loudly when the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("qcxms_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.qcxms_b200_egrad.argtypes = [C.c_int, ip, dp, C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, dp, ip]
        L.qcxms_b200_egrad_batch.argtypes = [C.c_int, C.c_int, ip, dp, C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, dp, ip, ip]
        L.qcxms_b200_fragment_structure.argtypes = [C.c_int, C.c_int, ip, dp, C.c_double, ip]
        L.qcxms_b200_ensemble_create.argtypes = [C.POINTER(MdConfig), C.c_int, C.c_int, ip, dp, C.c_int, C.POINTER(C.c_void_p)]
        L.qcxms_b200_ensemble_destroy.argtypes = [C.c_void_p]
        L.qcxms_b200_ensemble_set_trajectory.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, C.c_double, C.c_double]
        L.qcxms_b200_ensemble_set_all.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
        L.qcxms_b200_ensemble_run_md.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.qcxms_b200_basis_size.argtypes = [C.c_int, ip, C.c_int, C.POINTER(C.c_int32)]
        L.qcxms_b200_egrad_spec.argtypes = [C.c_int, ip, dp, C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, dp, C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32), dp, dp, dp]
        L.qcxms_b200_ensemble_set_warm_start.argtypes = [C.c_void_p, C.c_int]
        L.qcxms_b200_ensemble_set_mfp.argtypes = [C.c_void_p, C.c_int, dp]
        L.qcxms_b200_ensemble_get_new_velo.argtypes = [C.c_void_p, dp]
        L.qcxms_b200_ensemble_get_result.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, ip, dp, dp, C.POINTER(MdResult)]
        L.qcxms_b200_ensemble_get_all.argtypes = [C.c_void_p, dp, dp, dp, ip, dp, dp, C.POINTER(MdResult)]
        L.qcxms_b200_ensemble_last_timing.argtypes = [C.c_void_p, dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.qcxms_b200_ensemble_histogram.argtypes = [C.c_void_p, C.c_int, dp, C.POINTER(C.c_void_p)]
        L.qcxms_b200_ensemble_intenergy.argtypes = [C.c_void_p, dp, dp]
        L.qcxms_b200_cid_batch.argtypes = [C.POINTER(CidConfig), C.c_int, C.c_int, ip, dp, C.c_int, dp, dp, dp, dp, dp, ip, dp, dp, dp, ip,
                                           C.POINTER(CidResult), C.c_int]
        L.qcxms_b200_comm_unique_id.argtypes = [C.c_void_p]
        L.qcxms_b200_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.qcxms_b200_comm_destroy.argtypes = [C.c_void_p]
        L.qcxms_b200_comm_allreduce_sum.argtypes = [C.c_void_p, dp, C.c_int]
        L.qcxms_b200_ensemble_allreduce_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_int, dp]
        L.qcxms_b200_ensemble_set_esi.argtypes = [C.c_void_p, C.c_double]
        L.qcxms_b200_ensemble_set_gs_mode.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.qcxms_b200_ensemble_get_gs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp]
        L.qcxms_b200_last_error.restype = C.c_char_p
        L.qcxms_b200_version.restype = C.c_char_p
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _check(rc):
    if rc:
        raise RuntimeError("qcxms_b200 error %d: %s" % (rc, lib().qcxms_b200_last_error().decode()))


def version():
    return lib().qcxms_b200_version().decode()


def get_xtb_egrad(num, xyz, charge, multiplicity, method, etemp):
    """get_xtb_egrad(num, xyz, charge, multiplicity, method, etemp, ...) -> qat, energy, gradient, stat
    (reference src/tblite.f90:65-66; output_file / spec_calc are host-side concerns of the Fortran shim)."""
    num = np.ascontiguousarray(num, dtype=np.int32)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(len(num), 3)
    qat, grad = np.zeros(len(num)), np.zeros((len(num), 3))
    e, stat = C.c_double(0.0), C.c_int32(0)
    _check(lib().qcxms_b200_egrad(len(num), _ip(num), _dp(xyz), int(charge), int(multiplicity), int(method), float(etemp),
                                  _dp(qat), C.byref(e), _dp(grad), C.byref(stat)))
    return qat, e.value, grad, stat.value


def get_xtb_egrad_spec(num, xyz, charge, multiplicity, method, etemp, write_files=None):
    """get_xtb_egrad with spec_calc = .true. (reference src/tblite.f90:152-164): adds nao, ihomo, emo [nao] (Eh), focc [nao] and
    qmo [nao, nat] (write_qmo, src/mo_energ.f90:31-54).  write_files: directory to put tmp.mspec and qcxms.Mspec.tbxtb into, with the
    record structure of src/mo_energ.f90:56-74 (getspec reads them list-directed)."""
    num = np.ascontiguousarray(num, dtype=np.int32)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    nat = len(num)
    nao = C.c_int32(0)
    _check(lib().qcxms_b200_basis_size(nat, _ip(num), int(method), C.byref(nao)))
    n = nao.value
    qat = np.zeros(nat); grad = np.zeros((nat, 3)); e = C.c_double(0.0); stat = C.c_int32(0)
    emo, focc, qmo = np.zeros(n), np.zeros(n), np.zeros((n, nat))
    ihomo = C.c_int32(0)
    _check(lib().qcxms_b200_egrad_spec(nat, _ip(num), _dp(xyz), int(charge), int(multiplicity), int(method), float(etemp), _dp(qat),
                                       C.byref(e), _dp(grad), C.byref(stat), C.byref(nao), C.byref(ihomo), _dp(emo), _dp(focc), _dp(qmo)))
    out = dict(qat=qat, energy=e.value, gradient=grad, stat=stat.value, nao=n, ihomo=ihomo.value, emo=emo, focc=focc, qmo=qmo)
    if write_files is not None and stat.value == 0:
        with open(os.path.join(write_files, "qcxms.Mspec.tbxtb"), "w") as f:
            f.write(" %11d %11d\n" % (n, ihomo.value))
            for k in range(n):
                f.write("\n%3d %10.3f\n" % (k + 1, emo[k] * AUTOEV))
                f.write(" %6.2f\n" % focc[k])
                for j0 in range(0, nat, 10):
                    f.write("".join(" %6.2f" % (v * 100.0) for v in qmo[k, j0:j0 + 10]) + "\n")
        with open(os.path.join(write_files, "tmp.mspec"), "w") as f:
            f.write(" %11d %11d\n" % (n, ihomo.value))
            for k in range(n):
                f.write("  %.16E\n  %.16E\n" % (emo[k] * AUTOEV, focc[k]))
                for j in range(nat):
                    f.write("  %.16E\n" % qmo[k, j])
    return out


def egrad_batch(num, xyz, charge, multiplicity, method, etemp):
    """Batched get_xtb_egrad over xyz[nsys, nat, 3] -> dict(qat, energy, gradient, stat, niter)."""
    num = np.ascontiguousarray(num, dtype=np.int32)
    nat = len(num)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, nat, 3)
    nsys = xyz.shape[0]
    qat, grad, e = np.zeros((nsys, nat)), np.zeros((nsys, nat, 3)), np.zeros(nsys)
    stat, niter = np.zeros(nsys, dtype=np.int32), np.zeros(nsys, dtype=np.int32)
    _check(lib().qcxms_b200_egrad_batch(nsys, nat, _ip(num), _dp(xyz), int(charge), int(multiplicity), int(method), float(etemp),
                                        _dp(qat), _dp(e), _dp(grad), _ip(stat), _ip(niter)))
    return dict(qat=qat, energy=e, gradient=grad, stat=stat, niter=niter)


def fragment_structure(num, xyz, rcut=3.0):
    """fragment_structure(nat, oz, xyz, rcut, 1, 0, frag) for xyz[nsys, nat, 3] -> frag[nsys, nat] (int32)."""
    num = np.ascontiguousarray(num, dtype=np.int32)
    nat = len(num)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, nat, 3)
    frag = np.zeros((xyz.shape[0], nat), dtype=np.int32)
    _check(lib().qcxms_b200_fragment_structure(xyz.shape[0], nat, _ip(num), _dp(xyz), float(rcut), _ip(frag)))
    return frag


class Ensemble:
    """A batch of EI trajectories of one molecule: the replacement for one `qcxms --prod` process per
    TMPQCXMS/TMP.<n> directory (reference bin/pqcxms:88-98) running md() (reference src/md.f90:34)."""

    def __init__(self, num, mass, ntraj, mchrg=1, tstep_fs=0.5, nmax=10000, nfragexit=3, exit_rules=True, method=gfn2_xtb,
                 etemp=-1.0, ieetemp=0.0, ax=0.0, isec=1, device=0):
        self.num = np.ascontiguousarray(num, dtype=np.int32)
        self.nat = len(self.num)
        self.ntraj = int(ntraj)
        self.mass = np.ascontiguousarray(mass, dtype=np.float64)
        self.cfg = MdConfig(int(method), int(mchrg), int(nfragexit), int(bool(exit_rules)), int(nmax), int(isec),
                            float(tstep_fs) * FSTOAU, float(etemp), float(ieetemp), float(ax))
        self._h = C.c_void_p()
        _check(lib().qcxms_b200_ensemble_create(C.byref(self.cfg), self.ntraj, self.nat, _ip(self.num), _dp(self.mass), int(device),
                                                C.byref(self._h)))

    def close(self):
        if self._h:
            lib().qcxms_b200_ensemble_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_trajectory(self, itrj, xyz, velo, velof, eimp, tadd):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64); velo = np.ascontiguousarray(velo, dtype=np.float64)
        velof = np.ascontiguousarray(velof, dtype=np.float64)
        _check(lib().qcxms_b200_ensemble_set_trajectory(self._h, int(itrj), _dp(xyz), _dp(velo), _dp(velof), float(eimp), float(tadd)))

    def set_all(self, xyz, velo, velof, eimp, tadd):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (xyz, velo, velof, eimp, tadd)]
        assert a[0].size == self.ntraj * self.nat * 3 and a[3].size == self.ntraj
        _check(lib().qcxms_b200_ensemble_set_all(self._h, *[_dp(v) for v in a]))

    def set_warm_start(self, on=True):
        """Opt-in fast mode (not the reference protocol): SCC of each step starts from the previous step's converged populations."""
        _check(lib().qcxms_b200_ensemble_set_warm_start(self._h, int(bool(on))))

    def set_esi(self, tscale):
        """Heating MD before the first collision of an ESI/CID run (reference md() with method 3, icoll 0, starting_md): Berendsen scaling
        towards tscale (K) during the first nadd steps; eimp = E_Scale and tadd = pretadd as given to set_all."""
        _check(lib().qcxms_b200_ensemble_set_esi(self._h, float(tscale)))

    def set_gs_mode(self, it, tsoll=0.0):
        """md() with it = -1 (ground-state equilibration towards tsoll K) or it = 0 (NVE sampling, every step recorded); 1: production."""
        _check(lib().qcxms_b200_ensemble_set_gs_mode(self._h, int(it), float(tsoll)))

    def gs_records(self, itrj=0, first=0, count=None):
        """The records of qcxms.gs for trajectory itrj: [count, nat, 6] = xyz | velo of every step of the sampling run."""
        count = self.cfg.nmax - first if count is None else count
        out = np.zeros((count, self.nat, 6))
        _check(lib().qcxms_b200_ensemble_get_gs(self._h, int(itrj), int(first), int(count), _dp(out)))
        return out

    def set_mfp(self, icoll, new_velo):
        """Mean-free-path MD of a CID run: md() with the reference's method == 3, icoll >= 1; new_velo [ntraj] in m/s as cid() returned it."""
        nv = np.ascontiguousarray(new_velo, dtype=np.float64)
        assert nv.size == self.ntraj
        _check(lib().qcxms_b200_ensemble_set_mfp(self._h, int(icoll), _dp(nv)))

    def new_velo(self):
        nv = np.zeros(self.ntraj)
        _check(lib().qcxms_b200_ensemble_get_new_velo(self._h, _dp(nv)))
        return nv

    def run_md(self, max_steps=0):
        """Runs md() for every trajectory; returns the number of trajectory-MD-steps executed."""
        n = C.c_int64(0)
        _check(lib().qcxms_b200_ensemble_run_md(self._h, int(max_steps), C.byref(n)))
        return n.value

    def result(self, itrj):
        nat = self.nat
        out = dict(xyz=np.zeros((nat, 3)), velo=np.zeros((nat, 3)), grad=np.zeros((nat, 3)), list=np.zeros(nat, dtype=np.int32),
                   achrg=np.zeros(nat), axyz=np.zeros((nat, 3)))
        res = MdResult()
        _check(lib().qcxms_b200_ensemble_get_result(self._h, int(itrj), _dp(out["xyz"]), _dp(out["velo"]), _dp(out["grad"]),
                                                    _ip(out["list"]), _dp(out["achrg"]), _dp(out["axyz"]), C.byref(res)))
        for k, _ in MdResult._fields_:
            out[k] = getattr(res, k)
        return out

    def results(self):
        """All trajectories at once: dict of arrays with a leading [ntraj] axis + one record array of md() outputs."""
        nt, nat = self.ntraj, self.nat
        out = dict(xyz=np.zeros((nt, nat, 3)), velo=np.zeros((nt, nat, 3)), grad=np.zeros((nt, nat, 3)), list=np.zeros((nt, nat), dtype=np.int32),
                   achrg=np.zeros((nt, nat)), axyz=np.zeros((nt, nat, 3)))
        res = (MdResult * nt)()
        _check(lib().qcxms_b200_ensemble_get_all(self._h, _dp(out["xyz"]), _dp(out["velo"]), _dp(out["grad"]), _ip(out["list"]),
                                                 _dp(out["achrg"]), _dp(out["axyz"]), res))
        rec = np.frombuffer(res, dtype=np.dtype([(k, np.int32) for k, t in MdResult._fields_ if t is C.c_int32] +
                                                [(k, np.float64) for k, t in MdResult._fields_ if t is C.c_double])).copy()
        for k in rec.dtype.names:
            out[k] = rec[k]
        return out

    def last_timing(self):
        ms, launches, scc = C.c_double(0), C.c_int64(0), C.c_int64(0)
        _check(lib().qcxms_b200_ensemble_last_timing(self._h, C.byref(ms), C.byref(launches), C.byref(scc)))
        return dict(kernel_ms=ms.value, launches=launches.value, scc_iterations=scc.value)

    def intenergy(self):
        """intenergy (reference src/md.f90:715-741) of the current state: (fragT[ntraj, 10] / K, E_int[ntraj, 10] / Eh)."""
        T = np.zeros((self.ntraj, 10)); e = np.zeros((self.ntraj, 10))
        _check(lib().qcxms_b200_ensemble_intenergy(self._h, _dp(T), _dp(e)))
        return T, e

    def histogram(self, nbins=512):
        bins = np.zeros(nbins)
        dev = C.c_void_p()
        _check(lib().qcxms_b200_ensemble_histogram(self._h, int(nbins), _dp(bins), C.byref(dev)))
        return bins, dev.value


def cid_config(mchrg=1, gas="ar", tstep_fs=0.5, elab=40.0, ecom=0.0, eexact=False, manual_dist=0, ntot=15000, etemp=0.0, method=gfn2_xtb):
    z, m, _ = GASES[gas.lower()]
    return CidConfig(int(method), int(mchrg), z, int(bool(eexact)), int(manual_dist), int(ntot), m * AMUTOAU, float(tstep_fs) * FSTOAU,
                     float(etemp), float(elab), float(ecom))


def cid(cfg, num, mass, icoll, xyz, velo, rnd, velo_cm=None, direc=None, collided=None, device=0):
    """One collision of cid() (reference src/cid.f90:24-28) for a batch of ions xyz[ntraj, nuc, 3].
    rnd[ntraj, 9] are the uniform random numbers the reference draws inside the call.  Returns a dict of arrays
    (xyz, velo, grad, achrg, axyz, list, direc, collided) + the per-trajectory result fields."""
    num = np.ascontiguousarray(num, dtype=np.int32)
    nuc = len(num)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    xyz = np.array(xyz, dtype=np.float64).reshape(-1, nuc, 3)
    nt = xyz.shape[0]
    velo = np.array(velo, dtype=np.float64).reshape(nt, nuc, 3)
    rnd = np.ascontiguousarray(rnd, dtype=np.float64).reshape(nt, 9)
    vcm = None if velo_cm is None else np.ascontiguousarray(velo_cm, dtype=np.float64).reshape(nt)
    direc = np.zeros((nt, 3)) if direc is None else np.array(direc, dtype=np.float64).reshape(nt, 3)
    collided = np.zeros(nt, dtype=np.int32) if collided is None else np.array(collided, dtype=np.int32).reshape(nt)
    grad, achrg, axyz = np.zeros((nt, nuc, 3)), np.zeros((nt, nuc)), np.zeros((nt, nuc, 3))
    lst = np.zeros((nt, nuc), dtype=np.int32)
    res = (CidResult * nt)()
    _check(lib().qcxms_b200_cid_batch(C.byref(cfg), nt, nuc, _ip(num), _dp(mass), int(icoll), _dp(xyz), _dp(velo), _dp(rnd),
                                      None if vcm is None else _dp(vcm), _dp(direc), _ip(collided), _dp(grad), _dp(achrg), _dp(axyz), _ip(lst),
                                      res, int(device)))
    out = dict(xyz=xyz, velo=velo, grad=grad, achrg=achrg, axyz=axyz, list=lst, direc=direc, collided=collided)
    for k, t in CidResult._fields_:
        if k != "direc":
            out[k] = np.array([getattr(r, k) for r in res])
    return out


class Comm:
    """The path's one collective through the C ABI (qcxms_b200_comm_*: NCCL over NVLink / NVSwitch, one process per GPU).
    Rank 0 draws `Comm.unique_id()` and hands the 128 bytes to the other ranks (any transport the host has); every rank then
    constructs Comm(id, nranks, rank, device)."""

    def __init__(self, unique_id, nranks, rank, device=0):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._c = C.c_void_p()
        _check(lib().qcxms_b200_comm_create(buf, int(nranks), int(rank), int(device), C.byref(self._c)))
        self.nranks, self.rank = int(nranks), int(rank)

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        _check(lib().qcxms_b200_comm_unique_id(buf))
        return bytes(buf.raw)

    def allreduce_sum(self, arr):
        """in-place sum over the ranks of a host float64 array (e.g. spectrum.spectrum(records))"""
        a = np.ascontiguousarray(arr, dtype=np.float64)
        _check(lib().qcxms_b200_comm_allreduce_sum(self._c, _dp(a), int(a.size)))
        return a

    def allreduce_histogram(self, ensemble, nbins=512):
        bins = np.zeros(int(nbins))
        _check(lib().qcxms_b200_ensemble_allreduce_histogram(ensemble._h, self._c, int(nbins), _dp(bins)))
        return bins

    def close(self):
        if self._c:
            lib().qcxms_b200_comm_destroy(self._c)
            self._c = C.c_void_p()


def load_molecule(name):
    """Benchmark/test input geometries (bohr): the reference's share/examples molecules + caffeine."""
    with open(os.path.join(_HERE, "data", "molecules.json")) as f:
        m = json.load(f)[name]
    return np.array(m["num"], dtype=np.int32), np.array(m["xyz"], dtype=np.float64), int(m["charge"])
