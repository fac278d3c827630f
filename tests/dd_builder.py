"""Builds flat full-depth matrix DDs (fdd_matdd tables) for tests, without the reference:
a dense 2^k x 2^k matrix on an arbitrary set of target qubits, identity on all other levels.
Also numpy reference implementations used as an independent second checker."""
from __future__ import annotations

import numpy as np

from flatdd_b200 import FlatDD

TERMINAL = -1


def gate_dd(n: int, targets, matrix: np.ndarray, tol: float = 0.0) -> FlatDD:
    """`targets[i]` is the qubit that bit i of the dense matrix index refers to (bit 0 = targets[0])."""
    targets = list(targets)
    k = len(targets)
    matrix = np.asarray(matrix, dtype=np.complex128)
    assert matrix.shape == (1 << k, 1 << k)
    pos = {q: i for i, q in enumerate(targets)}
    level, child, weight = [], [], []
    unique = {}
    lowest = min(targets)

    def ident_chain(lv: int) -> int:
        if lv < 0:
            return TERMINAL
        key = ("ident", lv)
        if key not in unique:
            below = ident_chain(lv - 1)
            unique[key] = len(level)
            level.append(lv)
            child.append([below, TERMINAL, TERMINAL, below])
            weight.append([[1.0, 0.0], [0.0, 0.0], [0.0, 0.0], [1.0, 0.0]])
        return unique[key]

    def make(lv: int, rows: tuple, cols: tuple):
        """DD of the sub-matrix with the dense bits of levels > lv already fixed (rows/cols are
        dicts frozen as tuples of (bitpos, value)). Returns (node_index or TERMINAL, is_zero)."""
        rfix, cfix = dict(rows), dict(cols)
        # sub-matrix selector
        free = [pos[q] for q in targets if q <= lv]
        ridx = sum(v << b for b, v in rfix.items())
        cidx = sum(v << b for b, v in cfix.items())
        if lv < lowest:
            # below the lowest target everything is identity: one shared chain with unit weights, the
            # matrix entry sits on the edge into it (weights live near the top, as in a normalised DD)
            val = matrix[ridx, cidx]
            return (ident_chain(lv), complex(val)) if abs(val) > tol else (TERMINAL, 0j)
        # zero test on the whole remaining block
        sub = matrix
        r_sel = [ridx + sum(((m >> i) & 1) << b for i, b in enumerate(free)) for m in range(1 << len(free))]
        c_sel = [cidx + sum(((m >> i) & 1) << b for i, b in enumerate(free)) for m in range(1 << len(free))]
        block = sub[np.ix_(r_sel, c_sel)]
        if not np.any(np.abs(block) > tol):
            return TERMINAL, 0j
        edges = []
        if lv in pos:
            b = pos[lv]
            for rb in range(2):
                for cb in range(2):
                    edges.append(make(lv - 1, tuple(sorted({**rfix, b: rb}.items())), tuple(sorted({**cfix, b: cb}.items()))))
        else:
            d = make(lv - 1, rows, cols)
            edges = [d, (TERMINAL, 0j), (TERMINAL, 0j), d]
        # weights: child sub-DDs carry their own scalar; keep it on the edge
        key = (lv, tuple((c, w) for c, w in edges))
        if key not in unique:
            unique[key] = len(level)
            level.append(lv)
            child.append([c if w != 0 else TERMINAL for c, w in edges])
            weight.append([[w.real, w.imag] for c, w in edges])
        return unique[key], 1 + 0j

    # children return (node, weight): for internal nodes weight is 1, for terminals the matrix entry
    root, w = make(n - 1, (), ())
    assert w != 0, "zero matrix"
    return FlatDD(n, 4, root, np.array([1.0, 0.0]), np.array(level, dtype=np.int32), np.array(child, dtype=np.int32),
                  np.array(weight, dtype=np.float64))


def random_unitary(k: int, rng: np.random.Generator) -> np.ndarray:
    a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def random_state(n: int, rng: np.random.Generator):
    re = rng.normal(size=1 << n)
    im = rng.normal(size=1 << n)
    nrm = np.sqrt(np.sum(re * re) + np.sum(im * im))
    return re / nrm, im / nrm


def apply_dense(n: int, targets, matrix: np.ndarray, psi: np.ndarray) -> np.ndarray:
    """numpy reference: apply a dense gate on `targets` to a 2^n state (qubit 0 = LSB)."""
    targets = list(targets)
    k = len(targets)
    t = psi.reshape([2] * n)  # axis a <-> qubit n-1-a
    axes = [n - 1 - q for q in reversed(targets)]  # most significant dense bit first
    m = np.asarray(matrix, dtype=np.complex128).reshape([2] * (2 * k))
    out = np.tensordot(m, t, axes=(list(range(k, 2 * k)), axes))
    out = np.moveaxis(out, list(range(k)), axes)
    return np.ascontiguousarray(out).reshape(-1)


def controlled(u: np.ndarray, n_controls: int) -> np.ndarray:
    """Dense matrix of a gate with `n_controls` controls on the LOW dense bits and u on the top bits."""
    k = int(np.log2(u.shape[0]))
    dim = 1 << (k + n_controls)
    m = np.eye(dim, dtype=np.complex128)
    mask = (1 << n_controls) - 1
    for r in range(1 << k):
        for c in range(1 << k):
            m[(r << n_controls) | mask, (c << n_controls) | mask] = u[r, c]
    return m


def kron_dd(n: int, factors: dict) -> FlatDD:
    """Tensor product of one-qubit matrices: factors[q] is the 2x2 matrix on qubit q (identity elsewhere).
    One node per level, all four successors point to the node of the level below."""
    level, child, weight = [], [], []
    for lv in range(n - 1, -1, -1):
        m = np.asarray(factors.get(lv, np.eye(2)), dtype=np.complex128)
        nxt = len(level) + 1 if lv > 0 else TERMINAL
        level.append(lv)
        child.append([nxt if m[r, c] != 0 else TERMINAL for r in range(2) for c in range(2)])
        weight.append([[m[r, c].real, m[r, c].imag] for r in range(2) for c in range(2)])
    return FlatDD(n, 4, 0, np.array([1.0, 0.0]), np.array(level, dtype=np.int32), np.array(child, dtype=np.int32),
                  np.array(weight, dtype=np.float64))
