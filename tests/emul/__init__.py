"""TEST INFRASTRUCTURE: ctypes driver for the host SIMT emulation of the HOP select kernels."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "time-opt-ilqr_b200", "csrc")
_SO = os.path.join(_HERE, "libhop_emul.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class SelectArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("T_min", C.c_int), ("T_max", C.c_int), ("jitter", C.c_double),
                ("max_tries", C.c_int), ("A_aug", _dp), ("B_aug", _dp), ("Q_aug", _dp), ("R_inv", _dp), ("z0", _dp),
                ("QT", _dp), ("rinv_step_stride", C.c_long), ("w_explicit", _dp), ("J_out", _dp), ("T_out", _ip), ("Jstar_out", _dp), ("status", _ip),
                ("E_pre", _dp), ("X_pre", _dp), ("pre_bad", _ip), ("no_diag_fastpath", C.c_int)]


class FusedArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("T_min", C.c_int), ("T_max", C.c_int), ("jitter", C.c_double),
                ("max_tries", C.c_int), ("A", _dp), ("Bm", _dp), ("a_resid", _dp), ("X", _dp), ("U", _dp), ("u_stride", C.c_long), ("xg", _dp),
                ("w", _dp), ("u_ref", _dp), ("Q", _dp), ("R", _dp), ("Qf", _dp), ("wrap_mask", C.c_uint),
                ("q_reg", C.c_double), ("rho_reg", C.c_double), ("mode", C.c_int), ("skip", _ip), ("J_out", _dp), ("T_out", _ip), ("Jstar_out", _dp),
                ("status", _ip)]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("simt_emul.cpp", "simt_emul.h", "emul_select.cpp", "emul_ddp.cpp")]
    srcs += [os.path.join(_CSRC, f) for f in ("hop_select_core.cuh", "hop_select_body.cuh", "hop_simt.cuh", "hop_mma.cuh",
                                              "hop_select_mma_body.cuh", "hop_select_pipe_body.cuh", "hop_select_scan_body.cuh", "hop_select_gpipe_body.cuh", "hop_select_tpp_body.cuh", "hop_select_epl_body.cuh", "hop_select_ref_body.cuh", "hop_ddp_core.cuh", "hop_ddp_mma.cuh", "hop_dynamics.cuh")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DHOP_HOST_EMUL", "-Wno-unknown-pragmas", "-I" + _HERE, "-I" + _CSRC, "-fPIC",
                               "-shared", "-o", _SO, os.path.join(_HERE, "simt_emul.cpp"),
                               os.path.join(_HERE, "emul_select.cpp"), os.path.join(_HERE, "emul_ddp.cpp")])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def select_generic(A_aug, B_aug, Q_aug, R_inv, z0, QT, T_min, T_max, w_explicit=None, jitter=1e-9, max_tries=8,
                   mma=False, scan=False, pipe=False, tpp=False, ref=False, fp32=False, no_diag=False):
    A_aug, B_aug, Q_aug, R_inv, z0, QT = map(_d, (A_aug, B_aug, Q_aug, R_inv, z0, QT))
    Bsz, N, d = A_aug.shape[:3]
    m = B_aug.shape[3]
    wexp = None if w_explicit is None else _d(w_explicit)
    J = np.full((Bsz, T_max), np.nan); T = np.zeros(Bsz, np.int32); Js = np.zeros(Bsz); st = np.zeros(Bsz, np.int32)
    a = SelectArgs(Bsz, N, T_min, T_max, jitter, max_tries, _p(A_aug), _p(B_aug), _p(Q_aug), _p(R_inv), _p(z0), _p(QT),
                   (m * m if R_inv.ndim == 4 else 0), _p(wexp), _p(J), T.ctypes.data_as(_ip), _p(Js), st.ctypes.data_as(_ip),
                   None, None, None, int(bool(no_diag)))
    if ref or fp32:
        rc = lib().emul_select_generic_ref(d, m, C.byref(a), int(bool(fp32)))
        assert rc == 0, f"emulated kernel failed rc={rc}"
        return J, T, Js, st
    fn = lib().emul_select_generic_tpp if tpp else lib().emul_select_generic_pipe if pipe else lib().emul_select_generic_scan if scan else (lib().emul_select_generic_mma if mma else lib().emul_select_generic)
    rc = fn(d, m, C.byref(a))
    assert rc == 0, f"emulated kernel failed rc={rc}"
    return J, T, Js, st


def select_fused(A, Bm, a_resid, X, U, xg, w, u_ref, Q, R, Qf, wrap_mask, T_min, T_max, q_reg=1e-9, rho_reg=1e-12,
                 jitter=1e-9, max_tries=8, mma=False, mode=0, epl=False, ref=False):
    A, Bm, X, U, xg, w, u_ref, Q, R, Qf = map(_d, (A, Bm, X, U, xg, w, u_ref, Q, R, Qf))
    ar = None if a_resid is None else _d(a_resid)
    Bsz, N, n = A.shape[:3]
    m = Bm.shape[3]
    J = np.full((Bsz, T_max), np.nan); T = np.zeros(Bsz, np.int32); Js = np.zeros(Bsz); st = np.zeros(Bsz, np.int32)
    a = FusedArgs(Bsz, N, T_min, T_max, jitter, max_tries, _p(A), _p(Bm), _p(ar), _p(X), _p(U),
                  0 if U.ndim == 2 else U.shape[1] * U.shape[2], _p(xg), _p(w), _p(u_ref),
                  _p(Q), _p(R), _p(Qf), wrap_mask, q_reg, rho_reg, mode, None, _p(J), T.ctypes.data_as(_ip), _p(Js),
                  st.ctypes.data_as(_ip))
    rc = (lib().emul_select_fused_ref if ref else lib().emul_select_fused_epl if epl else lib().emul_select_fused_mma if mma else lib().emul_select_fused)(n, m, C.byref(a))
    assert rc == 0, f"emulated kernel failed rc={rc}"
    return J, T, Js, st


def chol_inv_mma(A):
    """hop::mma::chol_inv (layout L, blocked Gauss-Jordan) on one matrix, through the emulator."""
    A = _d(A)
    d = A.shape[0]
    X = np.zeros((d, d)); st = C.c_int(0)
    rc = lib().emul_chol_inv_mma(d, _p(A), _p(X), C.byref(st))
    assert rc == 0, rc
    return X, st.value


def backward_pass(A, Bm, X, U, xg, u_ref, Q, R, Qf, w, wrap_mask, T, lm, variant):
    """solver.backward_pass_truncated of ONE (n, m) = (12, 4) instance through the emulated kernels (variant 2: one warp, the
    reference's summation order; 3: matrix products as DMMA fragments).  Returns (k [T, m], K [T, m, n], ok, code)."""
    A, Bm, X, U, xg, u_ref, Q, R, Qf = map(_d, (A, Bm, X, U, xg, u_ref, Q, R, Qf))
    N, n = A.shape[0], A.shape[1]
    m = Bm.shape[2]
    k = np.zeros((N, m)); K = np.zeros((N, m, n)); ok = C.c_int(0)
    rc = lib().emul_backward(n, m, int(variant), int(T), C.c_double(lm), _p(A), _p(Bm), _p(X), _p(U), _p(xg), _p(u_ref), _p(Q), _p(R),
                             _p(Qf), C.c_double(w), C.c_uint(wrap_mask), _p(k), _p(K), C.byref(ok))
    assert rc >= 0, f"emulated backward pass failed rc={rc}"
    return k[:T], K[:T], bool(ok.value), rc
