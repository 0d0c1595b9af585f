"""ctypes binding of oracle/_build/liboracle.so — the CPU oracle (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
EXE = os.path.join(ROOT, "oracle", "_build", "nbody_oracle")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "nbody")  # unmodified samples/nbody.cc, built by oracle/Makefile

MODE_STRICT, MODE_SQRT3 = 0, 1
KIND_Q1, KIND_Q2, KIND_Q3 = 1, 2, 3
_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_ubyte)


class OrcEvents(C.Structure):
    _fields_ = [("min_d2", C.c_double), ("argmin_step", C.c_int), ("hit_step", C.c_int), ("destroyed_step", C.c_int),
                ("cost", C.c_double), ("steps_done", C.c_int), ("n_reach", C.c_int), ("reach_step", C.c_int * 64)]


class OrcAnswer(C.Structure):
    _fields_ = [("min_dist", C.c_double), ("hit_time_step", C.c_int), ("gravity_device_id", C.c_int),
                ("missile_cost", C.c_double), ("argmin_step", C.c_int), ("n_devices", C.c_int),
                ("device_index", C.c_int * 64), ("reach_step", C.c_int * 64), ("q3_hit_step", C.c_int * 64),
                ("q3_cost", C.c_double * 64)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_run_steps.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _up, C.c_int, C.c_int, C.c_int]
        L.orc_trajectory.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _up, C.c_int, C.c_int,
                                     C.c_int, C.POINTER(OrcEvents)]
        L.orc_solve.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _up, C.c_int, C.c_int,
                                C.POINTER(OrcAnswer)]
        _lib = L
    return _lib


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _u(a):
    assert a.dtype == np.uint8
    return a.ctypes.data_as(_up)


def run_steps(mode, n, q, v, m, is_device, step_begin, step_end, nthreads=None):
    """In place on planar q, v: steps step_begin+1 .. step_end (nbody.cc run_step)."""
    nthreads = nthreads or os.cpu_count()
    rc = lib().orc_run_steps(mode, n, _d(q), _d(v), _d(m), _u(is_device), step_begin, step_end, nthreads)
    assert rc == 0


def trajectory(mode, kind, sysm, destroy_device=-1, n_steps=200000, nthreads=None):
    nthreads = nthreads or os.cpu_count()
    q, v = sysm.q.copy(), sysm.v.copy()
    ev = OrcEvents()
    rc = lib().orc_trajectory(mode, kind, sysm.n, sysm.planet, sysm.asteroid, _d(q), _d(v), _d(sysm.m),
                              _u(sysm.is_device), destroy_device, n_steps, nthreads, C.byref(ev))
    assert rc == 0
    return ev, q, v


def solve(mode, sysm, n_steps=200000, nthreads=None):
    nthreads = nthreads or os.cpu_count()
    ans = OrcAnswer()
    rc = lib().orc_solve(mode, sysm.n, sysm.planet, sysm.asteroid, _d(sysm.q), _d(sysm.v), _d(sysm.m),
                         _u(sysm.is_device), n_steps, nthreads, C.byref(ans))
    assert rc == 0
    return ans
