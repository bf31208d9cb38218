"""oracle/_ref: the reference's OWN stop predictor, compiled from the unmodified source where it lies.

TEST INFRASTRUCTURE ONLY.  `build()` compiles /root/reference/gp_predictor/src/gp_predictor.cpp (unmodified, read in
place, never copied) with g++ against the stand-in ROS / Eigen / generated-message headers of oracle/ref_stubs/ and the
C driver oracle/ref_driver.cpp into oracle/_ref/libgp_predictor_ref.so.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot; /root/reference does not exist there, so `build()` only runs where the source is present
and `available()` says whether the library can be used.

What this pins: oracle/stop_oracle.c (and through it the CUDA look-ahead kernel) against the reference's own
GPCallBack / llh_to_enu code - operator order, the H aliasing index, the UT, the Joseph form, the loop structure and
the published stop time all come from the reference's text, not from a restatement.  What it does not pin: Eigen's
and roscpp's own code (absent from this image; the stand-ins are documented in their headers), i.e. the rounding
order inside a matrix product - so comparisons against _ref are to 1e-12 relative, decisions exact.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libgp_predictor_ref.so")
REF_ROOT = os.environ.get("CNGP_REFERENCE_ROOT", "/root/reference")
_REF_SRC = os.path.join(REF_ROOT, "gp_predictor", "src", "gp_predictor.cpp")
_REF_INC = os.path.join(REF_ROOT, "gp_predictor", "include")
_STUBS = os.path.join(_HERE, "ref_stubs")
_DRIVER = os.path.join(_HERE, "ref_driver.cpp")

INIT_LLH = (0.693457963620326, -1.39498384275845, 334.993517334743)     # gp_predictor/config/init_params.yaml
INIT_ECEF = (859153.0153, -4836303.7266, 4055378.501)


def _newest_input() -> float:
    t = max(os.path.getmtime(_DRIVER), os.path.getmtime(_REF_SRC))
    for root, _, files in os.walk(_STUBS):
        for f in files:
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force: bool = False) -> str | None:
    """Compile oracle/_ref when the reference source is present; return the .so path (or None if it cannot exist)."""
    if not os.path.exists(_REF_SRC):
        return _SO if os.path.exists(_SO) else None
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < _newest_input():
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        # -ffp-contract=off: no FMA formed behind the source's back (a catkin build without -march flags has none);
        # -Dmain=...: the node's main() (gp_predictor.cpp:180-190) must not collide with the host process
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-Dmain=ref_gp_predictor_node_main", "-I", _STUBS, "-I", _REF_INC,
                               _REF_SRC, _DRIVER, "-o", _SO])
    return _SO


def available() -> bool:
    try:
        return build() is not None
    except (subprocess.CalledProcessError, OSError):
        return os.path.exists(_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _lib = C.CDLL(so)
        _lib.ref_gp_callback.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_double, C.c_double] + \
            [C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.c_int,
             C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ref_llh_to_enu.argtypes = [C.c_double] * 3 + [C.c_void_p] * 3
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def llh_to_enu(lat, lon, h, init_llh=INIT_LLH, init_ecef=INIT_ECEF):
    out = np.zeros(3)
    il, ie = np.asarray(init_llh, dtype=np.float64), np.asarray(init_ecef, dtype=np.float64)
    assert lib().ref_llh_to_enu(lat, lon, h, _p(il), _p(ie), _p(out)) == 0
    return out


def gp_callback(mean, sigma, P, Q, STM, Hvec, pos, init_llh=INIT_LLH, init_ecef=INIT_ECEF, clock_arrive=0.0,
                clock_later=None, want_trace=True):
    """One GpPredictor::GPCallBack of the reference.  Returns dict(triggered, i_stop, stop_cmd, step_stop, n_steps,
    xy_err, xy_trace, P, K, R) - i_stop recovered from the published stop time under the constant clock."""
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    M = mean.size
    arrs = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (P, Q, STM, Hvec, pos, init_llh, init_ecef)]
    assert [a.size for a in arrs] == [225, 225, 225, 60, 3, 3, 3]
    const_clock = clock_later is None
    if const_clock:
        clock_later = clock_arrive
    npub, nsteps = C.c_int(), C.c_int()
    cmd, xy = C.c_double(), C.c_double()
    trace = np.full(5 * M, np.nan)
    P_out, K_out, R_out = np.zeros(225), np.zeros(60), np.zeros(16)
    rc = lib().ref_gp_callback(_p(mean), _p(sigma), M, *[_p(a) for a in arrs], clock_arrive, clock_later,
                               C.byref(npub), C.byref(cmd), C.byref(xy), C.byref(nsteps), _p(trace), trace.size,
                               _p(P_out), _p(K_out), _p(R_out))
    assert rc == 0
    trig = npub.value > 0
    out = dict(triggered=trig, stop_cmd=cmd.value, n_steps=nsteps.value, xy_err=xy.value,
               step_stop=nsteps.value - 1 if trig else 5 * M, P=P_out.reshape(15, 15), K=K_out.reshape(15, 4),
               R=R_out.reshape(4, 4))
    # i/10.0 is published under a constant clock; without a trigger every odometry update was consumed
    out["i_stop"] = int(round(cmd.value * 10.0)) if (trig and const_clock) else (None if trig else M)
    if want_trace:
        out["xy_trace"] = trace[:nsteps.value].copy()
    return out
