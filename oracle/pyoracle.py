"""ctypes access to the CPU oracle (TEST INFRASTRUCTURE; see oracle/xtb_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package qcxms_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("xtb_oracle.c", "md_oracle.c", "xtb_oracle.h", "md_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


class Detail(C.Structure):
    _fields_ = [("nsh", C.c_int), ("nao", C.c_int), ("niter", C.c_int), ("converged", C.c_int)] + \
        [(k, C.c_double) for k in ("e_rep", "e_disp_atm", "e_disp_sc", "e_el", "e_es2", "e_es3", "e_aes", "e_ts")] + \
        [(k, C.POINTER(C.c_double)) for k in ("cn", "cn_d4", "overlap", "h0", "dipole", "quadrupole", "emo", "focc",
                                              "qsh", "dpat", "qpat", "e_iter", "coeff")] + \
        [("ao2at", C.POINTER(C.c_int32)), ("ihomo", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        _LIB.xtb_oracle_egrad.argtypes = [C.c_int, ip, dp, C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, dp,
                                          C.POINTER(Detail)]
        _LIB.xtb_oracle_egrad.restype = C.c_int
        _LIB.xtb_oracle_dims.argtypes = [C.c_int, ip, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _LIB.xtb_oracle_syev.argtypes = [C.c_int, dp, dp]
        _LIB.xtb_oracle_set_accuracy.argtypes = [C.c_double]
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def dims(num, method=2):
    num = np.ascontiguousarray(num, dtype=np.int32)
    nsh, nao = C.c_int(), C.c_int()
    st = lib().xtb_oracle_dims(len(num), _ip(num), method, C.byref(nsh), C.byref(nao))
    if st:
        raise ValueError("unsupported composition/method")
    return nsh.value, nao.value


def egrad(num, xyz, charge=0, multiplicity=1, method=2, etemp=300.0, detail=False):
    """Returns dict(energy, gradient[nat,3], qat[nat], stat, ...detail arrays)."""
    num = np.ascontiguousarray(num, dtype=np.int32)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    nat = len(num)
    qat = np.zeros(nat)
    grad = np.zeros((nat, 3))
    e = C.c_double()
    out = {}
    if detail:
        nsh, nao = dims(num, method)
        d = Detail()
        bufs = dict(cn=np.zeros(nat), cn_d4=np.zeros(nat), overlap=np.zeros((nao, nao)), h0=np.zeros((nao, nao)),
                    dipole=np.zeros((3, nao, nao)), quadrupole=np.zeros((6, nao, nao)), emo=np.zeros(nao),
                    focc=np.zeros(nao), qsh=np.zeros(nsh), dpat=np.zeros((nat, 3)), qpat=np.zeros((nat, 6)),
                    e_iter=np.zeros(250), coeff=np.zeros((nao, nao)))
        for k, v in bufs.items():
            setattr(d, k, _dp(v))
        ao2at = np.zeros(nao, dtype=np.int32)
        d.ao2at = _ip(ao2at)
        bufs["ao2at"] = ao2at
        st = lib().xtb_oracle_egrad(nat, _ip(num), _dp(xyz), charge, multiplicity, method, etemp, _dp(qat),
                                    C.byref(e), _dp(grad), C.byref(d))
        out.update(bufs)
        for k in ("nsh", "nao", "niter", "converged", "e_rep", "e_disp_atm", "e_disp_sc", "e_el", "e_es2", "e_es3",
                  "e_aes", "e_ts", "ihomo"):
            out[k] = getattr(d, k)
    else:
        st = lib().xtb_oracle_egrad(nat, _ip(num), _dp(xyz), charge, multiplicity, method, etemp, _dp(qat),
                                    C.byref(e), _dp(grad), None)
    out.update(energy=e.value, gradient=grad, qat=qat, stat=st)
    return out


def qmo(ao2at, coeff, overlap, nat):
    """write_qmo of the reference without the files (src/mo_energ.f90:31-54): qmo[nao, nat], every orbital normalised over the atoms."""
    nao = len(ao2at)
    out = np.zeros((nao, nat))
    f = lib().xtb_oracle_qmo
    f.restype = None
    f.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f(int(nat), nao, _ip(np.ascontiguousarray(ao2at, dtype=np.int32)), _dp(np.ascontiguousarray(coeff)), _dp(np.ascontiguousarray(overlap)), _dp(out))
    return out


def set_accuracy(acc):
    lib().xtb_oracle_set_accuracy(float(acc))


def syev(a):
    a = np.array(a, dtype=np.float64, order="C")
    n = a.shape[0]
    w = np.zeros(n)
    st = lib().xtb_oracle_syev(n, _dp(a), _dp(w))
    return w, a, st


# ------------------------------------------------------------------ MD side (md_oracle.c)
class MdConfig(C.Structure):
    _fields_ = [("method_id", C.c_int32), ("mchrg", C.c_int32), ("nfragexit", C.c_int32), ("exit_rules", C.c_int32),
                ("nmax", C.c_int32), ("isec", C.c_int32), ("tstep", C.c_double), ("etemp_in", C.c_double),
                ("ieetemp", C.c_double), ("ax", C.c_double)]


class MdResult(C.Structure):
    _fields_ = [("mdok", C.c_int32), ("fragstate", C.c_int32), ("nstep", C.c_int32), ("nfrag", C.c_int32),
                ("status", C.c_int32), ("scc_iter_total", C.c_int32)] + \
        [(k, C.c_double) for k in ("Tav", "Epav", "Ekav", "aTlast", "dtime", "ttime", "Epot", "Ekin")]


def leapfrog(grad, mass, tstep, xyz, velo):
    xyz = np.array(xyz, dtype=np.float64); velo = np.array(velo, dtype=np.float64)
    grad = np.ascontiguousarray(grad, dtype=np.float64); mass = np.ascontiguousarray(mass, dtype=np.float64)
    ke = C.c_double()
    f = lib().md_oracle_leapfrog
    f.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f(len(mass), _dp(grad), _dp(mass), float(tstep), _dp(xyz), _dp(velo), C.byref(ke))
    return xyz, velo, ke.value


def ekinet(velo, mass):
    velo = np.ascontiguousarray(velo, dtype=np.float64); mass = np.ascontiguousarray(mass, dtype=np.float64)
    e, t = C.c_double(), C.c_double()
    f = lib().md_oracle_ekinet
    f.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f(len(mass), _dp(velo), _dp(mass), C.byref(e), C.byref(t))
    return e.value, t.value


def intenergy(lst, mass, velo, nfrag):
    """intenergy (reference src/md.f90:715-741): (T[10], E_int[10])."""
    lst = np.ascontiguousarray(lst, dtype=np.int32); mass = np.ascontiguousarray(mass, dtype=np.float64)
    velo = np.ascontiguousarray(velo, dtype=np.float64)
    T = np.zeros(10); e = np.zeros(10)
    f = lib().md_oracle_intenergy
    f.restype = None
    f.argtypes = [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f(len(mass), _ip(lst), _dp(mass), _dp(velo), int(nfrag), _dp(T), _dp(e))
    return T, e


def impactscale(velo, mass, velof, eimp, ff, e0):
    velo = np.array(velo, dtype=np.float64); mass = np.ascontiguousarray(mass, dtype=np.float64)
    velof = np.ascontiguousarray(velof, dtype=np.float64)
    f = lib().md_oracle_impactscale
    f.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double]
    err = f(len(mass), _dp(velo), _dp(mass), _dp(velof), float(eimp), float(ff), float(e0))
    return velo, err


def fragment_structure(oz, xyz, rcut=3.0, at1=1, at2=0):
    oz = np.ascontiguousarray(oz, dtype=np.int32); xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    frag = np.zeros(len(oz), dtype=np.int32)
    f = lib().md_oracle_fragment_structure
    f.argtypes = [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    f(len(oz), _ip(oz), _dp(xyz), float(rcut), at1, at2, _ip(frag))
    return frag


def fragmass(iat, list_, mass, imass=None):
    iat = np.ascontiguousarray(iat, dtype=np.int32); list_ = np.ascontiguousarray(list_, dtype=np.int32)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    nfrag = C.c_int32()
    fragx = np.zeros(10); fragat = np.zeros((10, 200), dtype=np.int32)
    f = lib().md_oracle_fragmass
    f.argtypes = [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                  C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    im = None if imass is None else _ip(np.ascontiguousarray(imass, dtype=np.int32))
    f(len(iat), _ip(iat), _ip(list_), _dp(mass), im, C.byref(nfrag), _dp(fragx), _ip(fragat))
    return nfrag.value, fragx, fragat


def checkqc(e, grad, qat, mchrg):
    grad = np.ascontiguousarray(grad, dtype=np.float64); qat = np.ascontiguousarray(qat, dtype=np.float64)
    ee = C.c_double(e)
    f = lib().md_oracle_checkqc
    f.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    ok = f(len(qat), C.byref(ee), _dp(grad), _dp(qat), int(mchrg))
    return bool(ok), ee.value


def setetemp(nfrag, eimp, ax=0.0, ieetemp=0.0):
    f = lib().md_oracle_setetemp
    f.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    f.restype = C.c_double
    return f(int(nfrag), float(eimp), float(ax), float(ieetemp))


def getspin(ic, chrg):
    ic = np.ascontiguousarray(ic, dtype=np.int32)
    f = lib().md_oracle_getspin
    f.argtypes = [C.c_int, C.POINTER(C.c_int32), C.c_int]
    return f(len(ic), _ip(ic), int(chrg))


def md(num, mass, xyz, velo, velof, eimp, tadd, mchrg=1, tstep_fs=0.5, nmax=10000, nfragexit=3, exit_rules=True, method=2,
       etemp=-1.0, ieetemp=0.0, ax=0.0, isec=1, max_steps=0):
    """md() of the reference for it > 0 (EI).  Returns dict with final xyz, velo, grad, list, achrg, axyz and result fields."""
    num = np.ascontiguousarray(num, dtype=np.int32); nat = len(num)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    xyz = np.array(xyz, dtype=np.float64).reshape(nat, 3); velo = np.array(velo, dtype=np.float64).reshape(nat, 3)
    velof = np.ascontiguousarray(velof, dtype=np.float64)
    cfg = MdConfig(int(method), int(mchrg), int(nfragexit), int(bool(exit_rules)), int(nmax), int(isec),
                   float(tstep_fs) * 41.3413733365614, float(etemp), float(ieetemp), float(ax))
    grad = np.zeros((nat, 3)); lst = np.zeros(nat, dtype=np.int32); achrg = np.zeros(nat); axyz = np.zeros((nat, 3))
    res = MdResult()
    f = lib().md_oracle_md
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    f.argtypes = [C.POINTER(MdConfig), C.c_int, ip, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int, dp, ip, dp, dp, C.POINTER(MdResult)]
    f(C.byref(cfg), nat, _ip(num), _dp(mass), _dp(xyz), _dp(velo), _dp(velof), float(eimp), float(tadd), int(max_steps),
      _dp(grad), _ip(lst), _dp(achrg), _dp(axyz), C.byref(res))
    out = dict(xyz=xyz, velo=velo, grad=grad, list=lst, achrg=achrg, axyz=axyz)
    for k, _ in MdResult._fields_:
        out[k] = getattr(res, k)
    return out


def md_esi(num, mass, xyz, velo, tscale, eimp, tadd, mchrg=1, tstep_fs=0.5, nmax=1000, method=2, etemp=-1.0, isec=1, exit_rules=True):
    """md() of the reference as the heating MD before the first collision of an ESI/CID run (method 3, icoll = 0, starting_md)."""
    num = np.ascontiguousarray(num, dtype=np.int32); nat = len(num)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    xyz = np.array(xyz, dtype=np.float64).reshape(nat, 3); velo = np.array(velo, dtype=np.float64).reshape(nat, 3)
    cfg = MdConfig(int(method), int(mchrg), 3, int(bool(exit_rules)), int(nmax), int(isec), float(tstep_fs) * 41.3413733365614, float(etemp), 0.0, 0.0)
    grad = np.zeros((nat, 3)); lst = np.zeros(nat, dtype=np.int32); achrg = np.zeros(nat); axyz = np.zeros((nat, 3))
    res = MdResult()
    f = lib().md_oracle_md_esi
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    f.argtypes = [C.POINTER(MdConfig), C.c_int, ip, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, dp, ip, dp, dp, C.POINTER(MdResult)]
    f(C.byref(cfg), nat, _ip(num), _dp(mass), _dp(xyz), _dp(velo), float(tscale), float(eimp), float(tadd), 0, _dp(grad), _ip(lst), _dp(achrg), _dp(axyz),
      C.byref(res))
    out = dict(xyz=xyz, velo=velo, grad=grad, list=lst, achrg=achrg, axyz=axyz)
    for k, _ in MdResult._fields_:
        out[k] = getattr(res, k)
    return out


def md_gs(num, mass, xyz, velo, it, tsoll, etemp, mchrg=0, tstep_fs=0.5, nmax=100, method=2, exit_rules=True):
    """md() of the reference for it = -1 (equilibration) / it = 0 (sampling).  Adds gs [nmax, nat, 6] (records of qcxms.gs) for it = 0."""
    num = np.ascontiguousarray(num, dtype=np.int32); nat = len(num)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    xyz = np.array(xyz, dtype=np.float64).reshape(nat, 3); velo = np.array(velo, dtype=np.float64).reshape(nat, 3)
    cfg = MdConfig(int(method), int(mchrg), 3, int(bool(exit_rules)), int(nmax), 1, float(tstep_fs) * 41.3413733365614, float(etemp), 0.0, 0.0)
    grad = np.zeros((nat, 3)); achrg = np.zeros(nat); gs = np.zeros((int(nmax), nat, 6))
    res = MdResult()
    f = lib().md_oracle_md_gs
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    f.argtypes = [C.POINTER(MdConfig), C.c_int, C.c_double, C.c_int, ip, dp, dp, dp, dp, dp, dp, C.POINTER(MdResult)]
    f(C.byref(cfg), int(it), float(tsoll), nat, _ip(num), _dp(mass), _dp(xyz), _dp(velo), _dp(grad), _dp(achrg), _dp(gs), C.byref(res))
    out = dict(xyz=xyz, velo=velo, grad=grad, achrg=achrg, gs=gs)
    for k, _ in MdResult._fields_:
        out[k] = getattr(res, k)
    return out


def md_mfp(num, mass, xyz, velo, icoll, new_velo, mchrg=1, tstep_fs=0.5, nmax=1000, method=2, etemp=-1.0, ieetemp=0.0, ax=0.0, isec=2,
           max_steps=0, exit_rules=True):
    """md() of the reference as the mean-free-path MD of a CID run (method 3, icoll >= 1).  Adds new_velo (m/s) to the result."""
    num = np.ascontiguousarray(num, dtype=np.int32); nat = len(num)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    xyz = np.array(xyz, dtype=np.float64).reshape(nat, 3); velo = np.array(velo, dtype=np.float64).reshape(nat, 3)
    cfg = MdConfig(int(method), int(mchrg), 3, int(bool(exit_rules)), int(nmax), int(isec),
                   float(tstep_fs) * 41.3413733365614, float(etemp), float(ieetemp), float(ax))
    grad = np.zeros((nat, 3)); lst = np.zeros(nat, dtype=np.int32); achrg = np.zeros(nat); axyz = np.zeros((nat, 3))
    res = MdResult()
    nv = C.c_double(float(new_velo))
    f = lib().md_oracle_md_mfp
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    f.argtypes = [C.POINTER(MdConfig), C.c_int, ip, dp, dp, dp, C.c_int, C.POINTER(C.c_double), C.c_int, dp, ip, dp, dp, C.POINTER(MdResult)]
    f(C.byref(cfg), nat, _ip(num), _dp(mass), _dp(xyz), _dp(velo), int(icoll), C.byref(nv), int(max_steps),
      _dp(grad), _ip(lst), _dp(achrg), _dp(axyz), C.byref(res))
    out = dict(xyz=xyz, velo=velo, grad=grad, list=lst, achrg=achrg, axyz=axyz, new_velo=nv.value)
    for k, _ in MdResult._fields_:
        out[k] = getattr(res, k)
    return out


# ------------------------------------------------------------------------------------------------ CID (md_oracle.c)
def eigvec3x3(a):
    a = np.array(a, dtype=np.float64).reshape(3, 3)
    w, q = np.zeros(3), np.zeros((3, 3))
    f = lib().md_oracle_eigvec3x3
    f.argtypes = [C.POINTER(C.c_double)] * 3
    f.restype = None
    f(_dp(a), _dp(w), _dp(q))
    return w, q


def euler_rotation(xyz, velo, a, b, c):
    xyz = np.array(xyz, dtype=np.float64); velo = np.array(velo, dtype=np.float64)
    f = lib().md_oracle_euler_rotation
    f.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double]
    f.restype = None
    f(len(xyz), _dp(xyz), _dp(velo), float(a), float(b), float(c))
    return xyz, velo


def rotation_velo(xyz, mass, velo):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); velo = np.ascontiguousarray(velo, dtype=np.float64)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    out, e = np.zeros_like(xyz), C.c_double(0.0)
    f = lib().md_oracle_rotation_velo
    f.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f.restype = None
    f(_dp(xyz), len(xyz), _dp(mass), _dp(velo), _dp(out), C.byref(e))
    return out, e.value


def vary_energies(e_in, e_distr, dum, dum2):
    f = lib().md_oracle_vary_energies
    f.argtypes = [C.c_double] * 4
    f.restype = C.c_double
    return f(float(e_in), float(e_distr), float(dum), float(dum2))


def cid(cfg, num, mass, icoll, xyz, velo, rnd, velo_cm=0.0, direc=None, collided=0):
    """cid() of the reference for one ion (mono-atomic gas); cfg is a qcxms_b200 CidConfig (same C layout)."""
    from qcxms_b200.api import CidResult
    num = np.ascontiguousarray(num, dtype=np.int32); nuc = len(num)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    xyz = np.array(xyz, dtype=np.float64).reshape(nuc, 3); velo = np.array(velo, dtype=np.float64).reshape(nuc, 3)
    rnd = np.ascontiguousarray(rnd, dtype=np.float64)
    direc = np.zeros(3) if direc is None else np.array(direc, dtype=np.float64)
    coll = C.c_int32(int(collided))
    grad, achrg, axyz, lst = np.zeros((nuc, 3)), np.zeros(nuc), np.zeros((nuc, 3)), np.zeros(nuc, dtype=np.int32)
    res = CidResult()
    f = lib().md_oracle_cid
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    f.argtypes = [C.c_void_p, C.c_int, ip, dp, C.c_int, dp, dp, dp, C.c_double, dp, ip, dp, dp, dp, ip, C.c_void_p]
    f(C.byref(cfg), nuc, _ip(num), _dp(mass), int(icoll), _dp(xyz), _dp(velo), _dp(rnd), float(velo_cm), _dp(direc), C.byref(coll),
      _dp(grad), _dp(achrg), _dp(axyz), _ip(lst), C.byref(res))
    out = dict(xyz=xyz, velo=velo, grad=grad, achrg=achrg, axyz=axyz, list=lst, direc=direc, collided=coll.value)
    for k, t in CidResult._fields_:
        if k != "direc":
            out[k] = getattr(res, k)
    return out


# ------------------------------------------------------------------------------------------------ fragment records
def boltz(temp, ip):
    ip = np.ascontiguousarray(ip, dtype=np.float64)
    pop = np.zeros_like(ip)
    f = lib().md_oracle_boltz
    f.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f.restype = None
    f(len(ip), float(temp), _dp(ip), _dp(pop))
    return pop


def res_line(charge, mchrg, itrj, isec, ifrag, pairs, icoll=None):
    types = np.array([p[0] for p in pairs], dtype=np.int32); counts = np.array([p[1] for p in pairs], dtype=np.int32)
    buf = C.create_string_buffer(256)
    f = lib().md_oracle_res_line
    f.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    f(buf, float(charge), int(mchrg), int(itrj), -1 if icoll is None else int(icoll), int(isec), int(ifrag), len(pairs), _ip(types), _ip(counts))
    return buf.value.decode()


def energies(jobs, etemp):
    """energy provider for qcxms_b200.fragments (CPU oracle instead of the CUDA batch entry point)"""
    from qcxms_b200.fragments import getspin
    es, st = [], []
    for num, xyz, chrg in jobs:
        r = egrad(num, xyz, charge=chrg, multiplicity=getspin(num, chrg), etemp=etemp)
        es.append(r["energy"]); st.append(r["stat"])
    return es, st


def md_batch(num, mass, xyz, velo, velof, eimp, tadd, mchrg, nmax, nfragexit, isec, tstep_fs, etemp):
    """md back end for qcxms_b200.production.run_ei (CPU oracle instead of the CUDA ensemble)"""
    return [md(num, mass, xyz[k], velo[k], velof[k], eimp[k], tadd[k], mchrg=mchrg, tstep_fs=tstep_fs, nmax=nmax, nfragexit=nfragexit,
               exit_rules=True, etemp=etemp, isec=isec) for k in range(len(xyz))]


def cid_batch(cfg, num, mass, icoll, xyz, velo, rnd, velo_cm, direc, collided):
    """cid back end for qcxms_b200.production.run_cid (CPU oracle instead of qcxms_b200_cid_batch)"""
    return [cid(cfg, num, mass, icoll, xyz[k], velo[k], rnd[k], velo_cm=velo_cm[k], direc=direc[k], collided=collided[k])
            for k in range(len(xyz))]


def esi_batch(num, mass, xyz, velo, tscale, e_scale, pretadd, mchrg, nmax, tstep_fs, etemp):
    """heating-MD back end for qcxms_b200.production.run_cid (CPU oracle instead of the CUDA ensemble)"""
    return [md_esi(num, mass, xyz[k], velo[k], tscale, e_scale, pretadd, mchrg=mchrg, tstep_fs=tstep_fs, nmax=nmax, etemp=etemp, isec=1)
            for k in range(len(xyz))]


def mfp_batch(num, mass, xyz, velo, new_velo, icoll, isec, mchrg, nmax, tstep_fs, etemp):
    """mean-free-path md back end for qcxms_b200.production.run_cid (CPU oracle instead of the CUDA ensemble)"""
    return [md_mfp(num, mass, xyz[k], velo[k], icoll, new_velo[k], mchrg=mchrg, tstep_fs=tstep_fs, nmax=nmax, etemp=etemp, isec=isec)
            for k in range(len(xyz))]
