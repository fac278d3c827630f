"""ctypes wrapper of oracle/libflat_oracle.so (the C restatement in flat_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs,
never by the flatdd_b200 package."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "libflat_oracle.so"
    src = _HERE / "flat_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fvisibility=hidden", "-fPIC", "-shared",
                        f"-I{_HERE.parent / 'include'}", str(src), "-o", str(so)], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        dp = ctypes.POINTER(ctypes.c_double)
        vp = ctypes.c_void_p
        L.oracle_convert.argtypes = [vp, dp, dp]
        L.oracle_convert_switch1.argtypes = [vp, ctypes.c_uint, dp, dp]
        L.oracle_dmavm.argtypes = [vp, dp, dp, dp, dp, ctypes.c_int]
        L.oracle_mac_count.argtypes = [vp]
        L.oracle_mac_count.restype = ctypes.c_uint64
        L.oracle_cost_ip.argtypes = [vp, ctypes.c_uint]
        L.oracle_cost_ip.restype = ctypes.c_uint64
        L.oracle_cost_op1.argtypes = [vp, ctypes.c_uint]
        L.oracle_cost_op1.restype = ctypes.c_uint64
        L.oracle_dd_size.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
        L.oracle_dd_size.restype = ctypes.c_uint64
        L.oracle_switch_index.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, dp, ctypes.c_int]
        _LIB = L
    return _LIB


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def convert(dd):
    """Exact per-amplitude walk (getValueByPathPar). Returns (re, im)."""
    dim = 1 << dd.n_qubits
    re = np.zeros(dim)
    im = np.zeros(dim)
    c = dd.as_c()
    lib().oracle_convert(ctypes.byref(c), _dp(re), _dp(im))
    return re, im


def convert_switch1(dd, n_thread_exp: int):
    """getVectorFromDDSwitch1 including the initial-regularity shortcut."""
    dim = 1 << dd.n_qubits
    re = np.zeros(dim)
    im = np.zeros(dim)
    c = dd.as_c()
    lib().oracle_convert_switch1(ctypes.byref(c), n_thread_exp, _dp(re), _dp(im))
    return re, im


def dmavm(gate, y_re: np.ndarray, y_im: np.ndarray):
    """DDArrMultiplyIP: returns z = M y as (re, im)."""
    y_re = np.ascontiguousarray(y_re, dtype=np.float64)
    y_im = np.ascontiguousarray(y_im, dtype=np.float64)
    z_re = np.zeros_like(y_re)
    z_im = np.zeros_like(y_im)
    c = gate.as_c()
    lib().oracle_dmavm(ctypes.byref(c), _dp(y_re), _dp(y_im), _dp(z_re), _dp(z_im), 0)
    return z_re, z_im


def mac_count(gate) -> int:
    c = gate.as_c()
    return int(lib().oracle_mac_count(ctypes.byref(c)))


def cost_ip(gate, n_thread_exp: int) -> int:
    c = gate.as_c()
    return int(lib().oracle_cost_ip(ctypes.byref(c), n_thread_exp))


def cost_op1(gate, n_thread_exp: int) -> int:
    c = gate.as_c()
    return int(lib().oracle_cost_op1(ctypes.byref(c), n_thread_exp))


def dd_size(dd) -> int:
    child = np.ascontiguousarray(dd.child, dtype=np.int32)
    return int(lib().oracle_dd_size(dd.n_nodes, dd.root, child.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), dd.radix))


def switch_index(n_qubits: int, sizes, beta: float = 0.9, threshold: float = 2.0) -> int:
    s = np.ascontiguousarray(sizes, dtype=np.float64)
    return int(lib().oracle_switch_index(n_qubits, beta, threshold, _dp(s), s.size))


def replay_trace(records):
    """Runs a boundary trace (convert, then every gate) on the CPU oracle. Returns (re, im)."""
    re = im = None
    for rec in records:
        if rec.kind == 1:
            re, im = convert(rec.dd)
        else:
            re, im = dmavm(rec.dd, re, im)
    return re, im
