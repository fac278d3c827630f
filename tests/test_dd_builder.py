"""CPU tests of the test-side DD builder and a second, independent check of the oracle:
oracle DMAVM on builder gates == numpy tensordot application."""
import numpy as np
import pytest

from oracle import pyoracle
from tests import dd_builder as B


@pytest.mark.parametrize("n,targets", [(3, [0]), (3, [2]), (4, [1, 3]), (5, [4, 0]), (6, [2, 5, 0]), (7, [6, 5])])
def test_builder_dense_matches_kron(n, targets):
    rng = np.random.default_rng(n * 17 + len(targets))
    u = B.random_unitary(len(targets), rng)
    dd = B.gate_dd(n, targets, u)
    dense = dd.to_dense()
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    assert np.allclose(dense @ psi, B.apply_dense(n, targets, u, psi), atol=1e-12)
    assert np.allclose(dense.conj().T @ dense, np.eye(1 << n), atol=1e-12)


@pytest.mark.parametrize("n,targets,kind", [
    (8, [0], "dense"), (8, [7], "dense"), (9, [3, 4], "dense"), (10, [0, 9], "dense"), (10, [2, 5, 7], "dense"),
    (11, [10, 4, 1], "ctrl"), (11, [0, 6, 9], "ctrl"), (12, [5, 6, 7, 8], "diag"), (12, [11, 0, 3], "perm"),
])
def test_oracle_dmavm_matches_numpy(n, targets, kind):
    rng = np.random.default_rng(n * 31 + sum(targets))
    k = len(targets)
    if kind == "dense":
        u = B.random_unitary(k, rng)
    elif kind == "ctrl":
        u = B.controlled(B.random_unitary(1, rng), k - 1)
    elif kind == "diag":
        u = np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << k)))
    else:
        u = np.eye(1 << k)[rng.permutation(1 << k)]
    dd = B.gate_dd(n, targets, u)
    yr, yi = B.random_state(n, rng)
    zr, zi = pyoracle.dmavm(dd, yr, yi)
    ref = B.apply_dense(n, targets, u, yr + 1j * yi)
    assert np.max(np.abs((zr + 1j * zi) - ref)) < 1e-14
    assert pyoracle.mac_count(dd) == int(np.count_nonzero(u)) << (n - k)
