"""CPU tests of the sharded (multi-GPU) path's host logic.

1. Single process: a boundary trace scheduled for G shards (exchange records + gates built in
   physical qubit order by the host driver) is replayed on a global-array model and must give the
   reference's final state -> validates the remap policy, the permuted gate DDs and the
   permutation bookkeeping, and that every gate is diagonal on the global qubits.
2. Two processes over gloo: the same trace replayed shard by shard (tests/cpu_shard_model.py),
   halves traded with send/recv like the library's NCCL path, then gathered and put back into
   logical order."""
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flatdd_b200 import read_trace
from flatdd_b200.sharded import replay, to_logical_order
from oracle import pyoracle
from tests import golden_util as G

ROOT = Path(__file__).resolve().parents[1]
SHARDED = ROOT / "tests" / "golden_sharded"
CASES = sorted(p.name for p in SHARDED.iterdir() if (p / "manifest.json").exists()) if SHARDED.exists() else []


def reference_final(case):
    m = json.loads((SHARDED / case / "manifest.json").read_text())
    stem = m["circuit"].replace(".qasm", "")
    fuse = m["trace_fuse"] if m["trace_fuse"] in (0, 1, 2) else 1
    golden = f"{stem}_f{fuse}"
    if golden not in G.cases():
        golden = f"{stem}_f1" if f"{stem}_f1" in G.cases() else f"{stem}_f0"
    fr, fi = G.final_state(golden)
    return m, fr + 1j * fi


class GlobalModel:
    """All shards in one array indexed by physical bits; an exchange is a swap of two index bits."""

    def __init__(self, n, n_local):
        self.n, self.n_local = n, n_local
        self.re = self.im = None

    def convert(self, dd):
        self.re, self.im = pyoracle.convert(dd)

    def apply(self, dd):
        # the gate must be diagonal on every global level (the kernels rely on it)
        for u in range(dd.n_nodes):
            if dd.level[u] >= self.n_local:
                assert not dd.weight[u, 1].any() and not dd.weight[u, 2].any(), "gate is non-diagonal on a global qubit"
        self.re, self.im = pyoracle.dmavm(dd, self.re, self.im)

    def exchange(self, pg, pl):
        assert pg >= self.n_local > pl >= 0
        idx = np.arange(1 << self.n)
        a, b = (idx >> pg) & 1, (idx >> pl) & 1
        src = np.where(a != b, idx ^ ((1 << pg) | (1 << pl)), idx)
        self.re, self.im = self.re[src], self.im[src]

    def relabel(self, a, b):
        pass


@pytest.mark.parametrize("case", CASES)
def test_sharded_schedule_on_global_model(case):
    m, want = reference_final(case)
    n, records = read_trace(SHARDED / case / "trace.bin")
    world = m["trace"]["world"]
    n_local = n - int(np.log2(world))
    model = GlobalModel(n, n_local)
    l2p = replay(records, model, n)
    got = to_logical_order(model.re + 1j * model.im, l2p)
    assert np.max(np.abs(got - want)) < 1e-12
    assert sum(r.kind == 3 for r in records) == m["trace"]["exchanges"]


class FusingModel(GlobalModel):
    """A backend that takes a gate and the exchange after it in one call, like GpuShard.apply_then_exchange (fdd_apply_many_exchange)."""

    def __init__(self, n, n_local):
        super().__init__(n, n_local)
        self.fused, self.plain_exchanges = 0, 0

    def exchange(self, pg, pl):
        self.plain_exchanges += 1
        super().exchange(pg, pl)

    def apply_then_exchange(self, dd, pg, pl):
        self.apply(dd)
        GlobalModel.exchange(self, pg, pl)
        self.fused += 1


@pytest.mark.parametrize("case", CASES)
def test_replay_hands_a_gate_and_the_exchange_after_it_to_one_call(case):
    """replay(): a gate record followed by an exchange record goes to apply_then_exchange when the backend has it; the layout map
    and the final state are those of the two separate calls, and every exchange is accounted for exactly once."""
    m, want = reference_final(case)
    n, records = read_trace(SHARDED / case / "trace.bin")
    n_local = n - int(np.log2(m["trace"]["world"]))
    plain, fusing = GlobalModel(n, n_local), FusingModel(n, n_local)
    l2p_plain = replay(records, plain, n)
    l2p = replay(records, fusing, n)
    assert l2p == l2p_plain
    assert np.array_equal(fusing.re, plain.re) and np.array_equal(fusing.im, plain.im)
    assert fusing.fused + fusing.plain_exchanges == sum(r.kind == 3 for r in records)
    follows_gate = sum(1 for a, b in zip(records, records[1:]) if a.kind == 2 and b.kind == 3)
    assert fusing.fused == follows_gate
    assert np.max(np.abs(to_logical_order(fusing.re + 1j * fusing.im, l2p) - want)) < 1e-12


def test_to_logical_order_roundtrip():
    rng = np.random.default_rng(0)
    n = 6
    psi = rng.normal(size=1 << n) + 0j
    l2p = [2, 0, 5, 1, 4, 3]
    phys = np.zeros_like(psi)
    for i in range(1 << n):
        j = sum(((i >> q) & 1) << l2p[q] for q in range(n))
        phys[j] = psi[i]
    assert np.array_equal(to_logical_order(phys, l2p), psi)


def _gloo_worker(rank, world, case, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.cpu_shard_model import CpuShard
        n, records = read_trace(SHARDED / case / "trace.bin")
        shard = CpuShard(n, rank, world)
        l2p = replay(records, shard, n)
        local = torch.from_numpy(np.stack([shard.re, shard.im]))
        gathered = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
        dist.gather(local, gathered, dst=0)
        if rank == 0:
            full = np.concatenate([g[0].numpy() + 1j * g[1].numpy() for g in gathered])  # rank = top physical bits
            np.save(Path(out_dir) / "state.npy", to_logical_order(full, l2p))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [c for c in CASES if c.endswith("_w2")])
def test_two_process_gloo_replay(case, tmp_path):
    m, want = reference_final(case)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, case, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "state.npy")
    assert np.max(np.abs(got - want)) < 1e-12
