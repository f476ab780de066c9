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
                                              "qsh", "dpat", "qpat", "e_iter")]


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
                    e_iter=np.zeros(250))
        for k, v in bufs.items():
            setattr(d, k, _dp(v))
        st = lib().xtb_oracle_egrad(nat, _ip(num), _dp(xyz), charge, multiplicity, method, etemp, _dp(qat),
                                    C.byref(e), _dp(grad), C.byref(d))
        out.update(bufs)
        for k in ("nsh", "nao", "niter", "converged", "e_rep", "e_disp_atm", "e_disp_sc", "e_el", "e_es2", "e_es3",
                  "e_aes", "e_ts"):
            out[k] = getattr(d, k)
    else:
        st = lib().xtb_oracle_egrad(nat, _ip(num), _dp(xyz), charge, multiplicity, method, etemp, _dp(qat),
                                    C.byref(e), _dp(grad), None)
    out.update(energy=e.value, gradient=grad, qat=qat, stat=st)
    return out


def set_accuracy(acc):
    lib().xtb_oracle_set_accuracy(float(acc))


def syev(a):
    a = np.array(a, dtype=np.float64, order="C")
    n = a.shape[0]
    w = np.zeros(n)
    st = lib().xtb_oracle_syev(n, _dp(a), _dp(w))
    return w, a, st
