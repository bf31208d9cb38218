"""ctypes wrapper around oracle/stop_oracle.c (the CPU restatement of GpPredictor::GPCallBack).

TEST INFRASTRUCTURE ONLY - see the header of stop_oracle.c.  Pinned on the reference's own code through oracle/_ref
(tests/test_ref_stop.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libstop_oracle.so")
_SRC = os.path.join(_HERE, "stop_oracle.c")


class StopCfg(C.Structure):
    _fields_ = [("v_nom", C.c_double), ("floor_a", C.c_double), ("floor_b", C.c_double), ("track", C.c_double),
                ("scale", C.c_double), ("thresh", C.c_double), ("ratio", C.c_int), ("fix_h_packing", C.c_int),
                ("trig_mode", C.c_int), ("pad_", C.c_int),
                ("init_llh", C.c_double * 3), ("init_ecef", C.c_double * 3)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.stop_oracle_default_cfg.argtypes = [C.POINTER(StopCfg)]
    return _lib


def default_cfg(**over) -> StopCfg:
    c = StopCfg()
    lib().stop_oracle_default_cfg(C.byref(c))
    for k, v in over.items():
        if k in ("init_llh", "init_ecef"):
            for j in range(3):
                getattr(c, k)[j] = float(v[j])
        else:
            setattr(c, k, v)
    return c


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def det_sincos(x):
    """The deterministic sin/cos shared bit for bit with the CUDA kernel (trig_mode = 1)."""
    s, c = C.c_double(), C.c_double()
    f = lib().stop_oracle_det_sincos
    f.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    f(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def llh_to_enu(lat, lon, h, cfg=None):
    cfg = cfg or default_cfg()
    out = np.zeros(3)
    f = lib().stop_oracle_llh_to_enu
    f.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(StopCfg), C.c_void_p]
    f(lat, lon, h, C.byref(cfg), _p(out))
    return out


def ut_R(mean, sigma, cfg=None):
    cfg = cfg or default_cfg()
    out = np.zeros((4, 4))
    f = lib().stop_oracle_ut_R
    f.argtypes = [C.c_double, C.c_double, C.POINTER(StopCfg), C.c_void_p]
    f(mean, sigma, C.byref(cfg), _p(out))
    return out


def lookahead(mean, sigma, P, Q, STM, Hvec, pos, cfg=None, want_trace=False):
    """One window.  Returns dict(triggered, i_stop, step_stop, xy_err, P[, xy_trace])."""
    cfg = cfg or default_cfg()
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    M = mean.size
    arrs = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (P, Q, STM, Hvec, pos)]
    assert [a.size for a in arrs] == [225, 225, 225, 60, 3]
    trig, i_stop, step = C.c_int(), C.c_int(), C.c_int()
    xy = C.c_double()
    trace = np.full(cfg.ratio * M, np.nan) if want_trace else None
    P_out, K_out, R_out = np.zeros(225), np.zeros(60), np.zeros(16)
    f = lib().stop_oracle_lookahead_ex
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(StopCfg)] + \
                 [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(mean), _p(sigma), M, *[_p(a) for a in arrs], C.byref(cfg), C.byref(trig), C.byref(i_stop), C.byref(step),
      C.byref(xy), _p(trace) if want_trace else None, _p(P_out), _p(K_out), _p(R_out))
    out = dict(triggered=bool(trig.value), i_stop=i_stop.value, step_stop=step.value, xy_err=xy.value,
               P=P_out.reshape(15, 15), K=K_out.reshape(15, 4), R=R_out.reshape(4, 4))
    if want_trace:
        out["xy_trace"] = trace
    return out


def lookahead_batch(mean, sigma, P, Q, STM, Hvec, pos, cfg=None):
    """B windows; each context array is either shared (one entry) or per window (leading dim B)."""
    cfg = cfg or default_cfg()
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    B, M = mean.shape
    sizes = (225, 225, 225, 60, 3)
    arrs, mask = [], 0
    for bit, (a, sz) in enumerate(zip((P, Q, STM, Hvec, pos), sizes)):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        if a.size == sz * B and B > 1:
            mask |= 1 << bit
        else:
            assert a.size == sz, (bit, a.size)
        arrs.append(a)
    trig = np.zeros(B, dtype=np.int32)
    i_stop = np.zeros(B, dtype=np.int32)
    step = np.zeros(B, dtype=np.int32)
    xy = np.zeros(B)
    f = lib().stop_oracle_lookahead_batch
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(StopCfg)] + \
                 [C.c_void_p] * 4
    f(_p(mean), _p(sigma), B, M, *[_p(a) for a in arrs], mask, C.byref(cfg), _p(trig), _p(i_stop), _p(step), _p(xy))
    return dict(triggered=trig, i_stop=i_stop, step_stop=step, xy_err=xy)
