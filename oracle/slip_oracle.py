"""ctypes wrapper around oracle/slip_oracle.c (CPU restatement of the slip extraction + GP_Input recorder of
CoreNav::Update, core_navigation/src/CoreNav.cpp:176-183, 190, 244-329).

TEST INFRASTRUCTURE ONLY - see the header of slip_oracle.c.  PARITY UNPINNED (no reference test exists)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libslip_oracle.so")
_SRC = os.path.join(_HERE, "slip_oracle.c")


class SlipCfg(C.Structure):
    _fields_ = [("wheel_radius", C.c_double), ("cmd_min", C.c_double), ("rear_min", C.c_double),
                ("arm_delay", C.c_int32), ("window", C.c_int32), ("min_samples", C.c_int32)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.slip_oracle_slip.restype = C.c_double
    return _lib


def default_cfg(**over) -> SlipCfg:
    c = SlipCfg()
    lib().slip_oracle_default_cfg(C.byref(c))
    for k, v in over.items():
        setattr(c, k, v)
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def slip_record(joint, att, vel, cmd, stop_cmd=None, max_windows=2, cap=149, cfg=None):
    """Same contract as GpContext.slip_record (host arrays)."""
    cfg = cfg or default_cfg()
    joint, att, vel, cmd = (np.ascontiguousarray(a, dtype=np.float64) for a in (joint, att, vel, cmd))
    if stop_cmd is not None:
        stop_cmd = np.ascontiguousarray(stop_cmd, dtype=np.float64)
    B, T = cmd.shape
    out = dict(slip=np.zeros((B, T)), time_array=np.zeros((B, max_windows, cap)), slip_array=np.zeros((B, max_windows, cap)),
               n_samples=np.zeros((B, max_windows), np.int32), published=np.zeros((B, max_windows), np.int32),
               stop_update=np.full((B, max_windows), -1, np.int32), n_windows=np.zeros(B, np.int32))
    lib().slip_record_batch(_p(joint), _p(att), _p(vel), _p(cmd), _p(stop_cmd), C.c_int64(B), C.c_int32(T), C.byref(cfg),
                            C.c_int32(max_windows), C.c_int32(cap), _p(out["slip"]), _p(out["time_array"]),
                            _p(out["slip_array"]), _p(out["n_samples"]), _p(out["published"]), _p(out["stop_update"]),
                            _p(out["n_windows"]))
    return out
