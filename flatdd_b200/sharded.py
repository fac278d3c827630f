"""Replay of a boundary trace on a sharded state (one process per shard) and the bookkeeping that
goes with it: the logical->physical qubit map, and putting a gathered state back into logical
order.  The same driver runs the GPU shards (GpuShard -> the C-ABI) and the CPU model the
gloo tests use (tests/cpu_shard_model.py); nothing here computes amplitudes."""
from __future__ import annotations

from typing import Iterable, List, Protocol

import numpy as np

from .flat import FlatDD, TraceRecord


class ShardBackend(Protocol):
    def convert(self, dd: FlatDD) -> None: ...
    def apply(self, dd: FlatDD) -> None: ...
    def exchange(self, global_physical_bit: int, local_physical_bit: int) -> None: ...
    def relabel(self, physical_bit_a: int, physical_bit_b: int) -> None: ...


def replay(records: Iterable[TraceRecord], backend: ShardBackend, n_qubits: int) -> List[int]:
    """Runs every record on the backend; returns the final logical->physical qubit map."""
    phys_to_logical = list(range(n_qubits))
    records = list(records)
    fuse = getattr(backend, "apply_then_exchange", None)  # a gate and the exchange that follows it in one boundary call
    skip = False
    for at, rec in enumerate(records):
        if skip:  # the exchange went along with the gate before it: only the map is left to do
            skip = False
            a, b = rec.exchange
            phys_to_logical[a], phys_to_logical[b] = phys_to_logical[b], phys_to_logical[a]
            continue
        if rec.kind == 1:
            backend.convert(rec.dd)
            phys_to_logical = list(range(n_qubits))
        elif rec.kind == 2:
            if fuse is not None and at + 1 < len(records) and records[at + 1].kind == 3:
                fuse(rec.dd, *records[at + 1].exchange)
                skip = True
            else:
                backend.apply(rec.dd)
        elif rec.kind in (3, 4):
            a, b = rec.exchange
            if rec.kind == 3:
                backend.exchange(a, b)
            else:
                backend.relabel(a, b)
            phys_to_logical[a], phys_to_logical[b] = phys_to_logical[b], phys_to_logical[a]
        else:
            raise ValueError(f"unknown trace record kind {rec.kind}")
    logical_to_physical = [0] * n_qubits
    for p, q in enumerate(phys_to_logical):
        logical_to_physical[q] = p
    return logical_to_physical


def to_logical_order(state: np.ndarray, logical_to_physical: List[int]) -> np.ndarray:
    """state is indexed by physical bits (bit p of the index = physical qubit p); returns the array
    indexed by logical bits.  Plain axis permutation, for test-sized states."""
    n = len(logical_to_physical)
    t = state.reshape([2] * n)  # axis a <-> physical bit n-1-a
    # logical bit q lives on physical bit l2p[q]: new axis (n-1-q) = old axis (n-1-l2p[q])
    axes = [n - 1 - logical_to_physical[n - 1 - a] for a in range(n)]
    return np.ascontiguousarray(np.transpose(t, axes)).reshape(-1)


class GpuShard:
    """ShardBackend on top of a sharded Context (fdd_create_sharded + fdd_comm_init)."""

    def __init__(self, ctx, exchange_method: int = 0):
        self.ctx = ctx
        self.method = exchange_method
        self.exchanges = 0

    def convert(self, dd):
        self.ctx.convert(dd)

    def apply(self, dd):
        self.ctx.apply(dd)

    def exchange(self, g, l):
        self.ctx.exchange_qubits(g, l, self.method)
        self.exchanges += 1

    def apply_then_exchange(self, dd, g, l):
        if self.method != 0:  # the fused exchange is the peer-memory path
            self.apply(dd)
            self.exchange(g, l)
            return
        self.ctx.apply_many_exchange([dd], g, l)
        self.exchanges += 1

    def relabel(self, a, b):
        self.ctx.relabel_qubits(a, b)
