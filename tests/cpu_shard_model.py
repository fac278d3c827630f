"""CPU model of one shard of a distributed state, for the world_size-2 gloo tests: numpy arrays,
the C oracle for the shard-local arithmetic and torch.distributed (gloo) for the half-shard
exchange.  It follows the protocol of the library's NCCL path (flatdd_b200/csrc/api.cu,
exchangeBits method 1): the half whose local bit equals the rank's global bit stays, the other
half is traded with the rank that differs in that global bit."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from flatdd_b200 import FlatDD
from oracle import pyoracle

TERMINAL = -1


def restrict_to_shard(dd: FlatDD, n_local: int, rank: int) -> FlatDD:
    """Sub-DD seen by shard `rank`: follow the successor selected by the rank's bits on the global
    levels (vector DD: e[bit]; matrix DD: the diagonal successor e[3*bit], the gate must be diagonal
    there), multiply the weights, and renumber the nodes below into an n_local-qubit table."""
    node = dd.root
    w = complex(*dd.root_weight)
    for lv in range(dd.n_qubits - 1, n_local - 1, -1):
        assert dd.level[node] == lv
        bit = (rank >> (lv - n_local)) & 1
        if dd.radix == 4:
            off = [1, 2] if bit == 0 else [1, 2]
            assert all(complex(*dd.weight[node, k]) == 0 for k in off), "gate is not diagonal on a global qubit"
            k = 3 * bit
        else:
            k = bit
        w *= complex(*dd.weight[node, k])
        node = int(dd.child[node, k])
        if w == 0:
            break
    if w == 0 or node == TERMINAL:
        # zero contribution: a table with a root whose edges are all zero
        lv = np.arange(n_local - 1, -1, -1, dtype=np.int32)
        child = np.full((n_local, dd.radix), TERMINAL, dtype=np.int32)
        return FlatDD(n_local, dd.radix, 0, np.zeros(2), lv, child, np.zeros((n_local, dd.radix, 2)))
    # renumber reachable nodes
    order, index = [node], {node: 0}
    for u in order:
        for k in range(dd.radix):
            c = int(dd.child[u, k])
            if c != TERMINAL and complex(*dd.weight[u, k]) != 0 and c not in index:
                index[c] = len(order)
                order.append(c)
    level = dd.level[order]
    child = np.full((len(order), dd.radix), TERMINAL, dtype=np.int32)
    weight = dd.weight[order].copy()
    for i, u in enumerate(order):
        for k in range(dd.radix):
            c = int(dd.child[u, k])
            if c != TERMINAL and complex(*dd.weight[u, k]) != 0:
                child[i, k] = index[c]
    return FlatDD(n_local, dd.radix, 0, np.array([w.real, w.imag]), level, child, weight)


class CpuShard:
    def __init__(self, n_qubits: int, rank: int, world: int):
        self.n = n_qubits
        self.rank = rank
        self.world = world
        self.n_local = n_qubits - int(np.log2(world))
        self.re = np.zeros(1 << self.n_local)
        self.im = np.zeros(1 << self.n_local)

    def convert(self, dd):
        self.re, self.im = pyoracle.convert(restrict_to_shard(dd, self.n_local, self.rank))

    def apply(self, dd):
        self.re, self.im = pyoracle.dmavm(restrict_to_shard(dd, self.n_local, self.rank), self.re, self.im)

    def relabel(self, a, b):
        pass

    def exchange(self, pg, pl):
        gbit = pg - self.n_local
        partner = self.rank ^ (1 << gbit)
        my_bit = (self.rank >> gbit) & 1
        idx = np.arange(1 << self.n_local)
        trade = ((idx >> pl) & 1) != my_bit  # this half leaves, the partner's matching half arrives
        send = torch.from_numpy(np.stack([self.re[trade], self.im[trade]]))
        recv = torch.empty_like(send)
        ops = [dist.P2POp(dist.isend, send, partner), dist.P2POp(dist.irecv, recv, partner)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.re[trade] = recv[0].numpy()
        self.im[trade] = recv[1].numpy()
