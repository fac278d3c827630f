"""Flat decision-diagram tables (numpy side of include/flatdd_b200.h) and the binary files
that carry them: single tables written by oracle/ref_dump.cpp and boundary traces written by
flatdd_b200/host/array_backend.hpp (TraceRecorder)."""
from __future__ import annotations

import ctypes
import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import List

import numpy as np

TERMINAL = -1


class _CVecDD(ctypes.Structure):
    _fields_ = [
        ("n_qubits", ctypes.c_int32),
        ("n_nodes", ctypes.c_int32),
        ("root", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("root_weight", ctypes.c_double * 2),
        ("level", ctypes.POINTER(ctypes.c_int32)),
        ("child", ctypes.POINTER(ctypes.c_int32)),
        ("weight", ctypes.POINTER(ctypes.c_double)),
    ]


# fdd_vecdd and fdd_matdd share one layout; only the radix of the tables differs
CVecDD = _CVecDD
CMatDD = _CVecDD


@dataclass
class FlatDD:
    """radix 2 = vector DD (fdd_vecdd), radix 4 = matrix DD (fdd_matdd, successor 2*row+col)."""

    n_qubits: int
    radix: int
    root: int
    root_weight: np.ndarray  # (2,) float64
    level: np.ndarray  # (n_nodes,) int32
    child: np.ndarray  # (n_nodes, radix) int32
    weight: np.ndarray  # (n_nodes, radix, 2) float64
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        self.root_weight = np.ascontiguousarray(self.root_weight, dtype=np.float64).reshape(2)
        self.level = np.ascontiguousarray(self.level, dtype=np.int32).reshape(-1)
        self.child = np.ascontiguousarray(self.child, dtype=np.int32).reshape(-1, self.radix)
        self.weight = np.ascontiguousarray(self.weight, dtype=np.float64).reshape(-1, self.radix, 2)

    @property
    def n_nodes(self) -> int:
        return int(self.level.shape[0])

    def as_c(self) -> _CVecDD:
        c = _CVecDD()
        c.n_qubits = self.n_qubits
        c.n_nodes = self.n_nodes
        c.root = self.root
        c.reserved = 0
        c.root_weight[0] = float(self.root_weight[0])
        c.root_weight[1] = float(self.root_weight[1])
        c.level = self.level.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        c.child = self.child.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        c.weight = self.weight.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        return c

    # -- dense views, for small test cases only ------------------------------------------------
    def to_dense(self) -> np.ndarray:
        """Vector of 2^n amplitudes (radix 2) or 2^n x 2^n matrix (radix 4), by plain recursion."""
        n = self.n_qubits
        memo = {}

        def sub(node: int, lv: int) -> np.ndarray:
            if node == TERMINAL:
                return np.ones((1,) if self.radix == 2 else (1, 1), dtype=np.complex128)
            if node in memo:
                return memo[node]
            half = 1 << lv
            if self.radix == 2:
                out = np.zeros(2 * half, dtype=np.complex128)
                for b in range(2):
                    w = complex(*self.weight[node, b])
                    if w != 0:
                        out[b * half:(b + 1) * half] = w * sub(int(self.child[node, b]), lv - 1)
            else:
                out = np.zeros((2 * half, 2 * half), dtype=np.complex128)
                for r in range(2):
                    for c in range(2):
                        w = complex(*self.weight[node, 2 * r + c])
                        if w != 0:
                            out[r * half:(r + 1) * half, c * half:(c + 1) * half] = w * sub(int(self.child[node, 2 * r + c]), lv - 1)
            memo[node] = out
            return out

        return complex(*self.root_weight) * sub(self.root, n - 1)

    def write(self, path) -> None:
        with open(path, "wb") as f:
            f.write(struct.pack("<4i", self.n_qubits, self.n_nodes, self.root, self.radix))
            f.write(self.root_weight.tobytes())
            f.write(self.level.tobytes())
            f.write(self.child.tobytes())
            f.write(self.weight.tobytes())


def _parse_tables(buf: memoryview, off: int, n_nodes: int, radix: int):
    level = np.frombuffer(buf, dtype="<i4", count=n_nodes, offset=off).copy()
    off += 4 * n_nodes
    child = np.frombuffer(buf, dtype="<i4", count=n_nodes * radix, offset=off).copy()
    off += 4 * n_nodes * radix
    weight = np.frombuffer(buf, dtype="<f8", count=n_nodes * radix * 2, offset=off).copy()
    off += 8 * n_nodes * radix * 2
    return level, child, weight, off


def read_flat(path) -> FlatDD:
    """Single table file: int32 n_qubits, n_nodes, root, radix; double root_weight[2]; tables."""
    buf = memoryview(Path(path).read_bytes())
    n_qubits, n_nodes, root, radix = struct.unpack_from("<4i", buf, 0)
    rw = np.frombuffer(buf, dtype="<f8", count=2, offset=16).copy()
    level, child, weight, _ = _parse_tables(buf, 32, n_nodes, radix)
    return FlatDD(n_qubits, radix, root, rw, level, child, weight)


@dataclass
class TraceRecord:
    kind: int  # 1 = convert, 2 = apply, 3 = exchange (global, local) physical bits, 4 = relabel two physical bits
    n_original_gates: int
    dd: FlatDD | None
    exchange: tuple | None = None  # (global physical bit, local physical bit) for kind 3


def read_trace(path) -> tuple[int, List[TraceRecord]]:
    """Boundary trace (TraceRecorder), raw or gzip'd (*.gz): returns (n_qubits, records)."""
    raw = Path(path).read_bytes()
    if raw[:2] == b"\x1f\x8b":
        import gzip
        raw = gzip.decompress(raw)
    buf = memoryview(raw)
    if bytes(buf[:8]) != b"FDDTRC01":
        raise ValueError(f"{path}: not a flatdd_b200 trace")
    n_qubits, n_records = struct.unpack_from("<2i", buf, 8)
    off = 16
    records: List[TraceRecord] = []
    for _ in range(n_records):
        kind, n_nodes, root, n_orig = struct.unpack_from("<4i", buf, off)
        off += 16
        rw = np.frombuffer(buf, dtype="<f8", count=2, offset=off).copy()
        off += 16
        if kind in (3, 4):  # exchange / relabel of two physical bits
            records.append(TraceRecord(kind, 0, None, (n_nodes, root)))
            continue
        if kind == 5:  # meta: world size the schedule was made for
            continue
        radix = 2 if kind == 1 else 4
        level, child, weight, off = _parse_tables(buf, off, n_nodes, radix)
        records.append(TraceRecord(kind, n_orig, FlatDD(n_qubits, radix, root, rw, level, child, weight)))
    return n_qubits, records
