"""Multi-GPU parity (needs >= 2 GPUs; `gpurun --gpus 2` or more).  One process per GPU: sharded
contexts over the C-ABI, NCCL unique id broadcast over gloo, boundary traces scheduled for the
shard count, both exchange methods (0 = peer-memory kernel, 1 = NCCL send/recv), results
gathered and compared with the reference's final state."""
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flatdd_b200 import read_trace
from flatdd_b200.sharded import GpuShard, replay, to_logical_order
from tests import golden_util as G
from tests.test_sharded_cpu import CASES, SHARDED, reference_final

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]


def n_gpus() -> int:
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worker(rank, world, trace_path, port, out_dir, method, canonicalize):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flatdd_b200 import Context, load_library
        lib = load_library()
        uid = [lib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        n, records = read_trace(trace_path)
        with Context(n, device=rank, rank=rank, world_size=world, library=lib) as ctx:
            ctx.comm_init(uid[0])
            shard = GpuShard(ctx, exchange_method=method)
            l2p = replay(records, shard, n)
            assert list(ctx.permutation()) == l2p  # the context mirrors the layout
            if rank == 0:
                (Path(out_dir) / "stats.json").write_text(json.dumps({"exchanges": ctx.get_option("exchanges"), "fused_exchanges": ctx.get_option("fused_exchanges")}))
            if canonicalize:
                ctx.canonicalize()
                assert list(ctx.permutation()) == list(range(n))
                l2p = list(range(n))
            re, im = ctx.get_state()
            ctx.barrier()
        local = torch.from_numpy(np.stack([re, im]))
        gathered = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
        dist.gather(local, gathered, dst=0)
        if rank == 0:
            full = np.concatenate([g[0].numpy() + 1j * g[1].numpy() for g in gathered])
            np.save(Path(out_dir) / "state.npy", to_logical_order(full, l2p))
    finally:
        dist.destroy_process_group()


def run_sharded(trace_path, world, tmp_path, method=0, canonicalize=False):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, str(trace_path), port, str(tmp_path), method, canonicalize), nprocs=world, join=True)
    return np.load(Path(tmp_path) / "state.npy")


def _world(case):
    return json.loads((SHARDED / case / "manifest.json").read_text())["trace"]["world"]


@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_sharded_trace_vs_reference(case, method, tmp_path):
    world = _world(case)
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _, want = reference_final(case)
    got = run_sharded(SHARDED / case / "trace.bin", world, tmp_path, method=method)
    assert np.max(np.abs(got - want)) < 1e-10
    assert 1.0 - abs(np.vdot(got, want)) ** 2 / (np.vdot(got, got).real * np.vdot(want, want).real) < 1e-10


def test_exchange_fused_into_the_preceding_pass(tmp_path):
    """fdd_apply_many_exchange: the pass before an exchange writes the traded half straight into the partner's buffer.  The
    replay hands every gate that is followed by an exchange to that call; on a dense-block schedule most exchanges then cost no
    pass of their own, and the state is the reference's."""
    trace = G.TRACES / "supremacy_n20_gpu_w2" / "trace.bin"
    if n_gpus() < 2 or not trace.exists() or "supremacy_n20_f1" not in G.cases(G.TRAVEL):
        pytest.skip("needs 2 GPUs and the travel goldens")
    got = run_sharded(trace, 2, tmp_path)
    stats = json.loads((Path(tmp_path) / "stats.json").read_text())
    assert stats["exchanges"] > 0 and stats["fused_exchanges"] > 0, stats
    if (G.TRAVEL / "supremacy_n20_f1" / "final_re.f64").exists():
        fr, fi = G.final_state("supremacy_n20_f1", G.TRAVEL)
        assert np.max(np.abs(got - (fr + 1j * fi))) < 1e-10
    else:
        idx, sr, si = G.samples("supremacy_n20_f1", G.TRAVEL)
        assert np.max(np.abs(got[idx.astype(np.int64)] - (sr + 1j * si))) < 1e-10


@pytest.mark.parametrize("case", [c for c in CASES if c.endswith("_w2")][:2])
def test_canonicalize_restores_logical_layout(case, tmp_path):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    _, want = reference_final(case)
    got = run_sharded(SHARDED / case / "trace.bin", 2, tmp_path, method=0, canonicalize=True)
    assert np.max(np.abs(got - want)) < 1e-10


@pytest.mark.parametrize("name,golden", [("supremacy_n20_gpu_w2", "supremacy_n20_f1"), ("knn_n25_f0_w2", "knn_n25_f1")])
def test_sharded_reference_circuits(name, golden, tmp_path):
    """Reference circuits sharded over 2 GPUs against golden data of the compiled reference."""
    trace = G.TRACES / name / "trace.bin"
    if n_gpus() < 2 or not trace.exists() or golden not in G.cases(G.TRAVEL):
        pytest.skip("needs 2 GPUs and the travel goldens")
    got = run_sharded(trace, 2, tmp_path)
    if (G.TRAVEL / golden / "final_re.f64").exists():
        fr, fi = G.final_state(golden, G.TRAVEL)
        assert np.max(np.abs(got - (fr + 1j * fi))) < 1e-10
    else:
        idx, sr, si = G.samples(golden, G.TRAVEL)
        assert np.max(np.abs(got[idx.astype(np.int64)] - (sr + 1j * si))) < 1e-10


def test_sharded_cli_two_processes(tmp_path):
    """The C++ host binary itself, one process per GPU: rendezvous through a file, same circuit on
    both ranks, shards written per rank and compared with the reference's final state."""
    import subprocess
    from tests.test_gpu_cli import CLI
    if n_gpus() < 2 or not CLI.exists():
        pytest.skip("needs 2 GPUs and build/flatdd_gpu")
    circuit = ROOT / "tests" / "circuits" / "mix_n12.qasm"
    cwd = tmp_path / "build" / "apps"
    cwd.mkdir(parents=True)
    (tmp_path / "log" / "results" / "time").mkdir(parents=True)
    (tmp_path / "log" / "results" / "state").mkdir(parents=True)
    procs = [subprocess.Popen([str(CLI), "--file", str(circuit), "-t", "8", "--fuse", "3", "--quiet", "--world", "2", "--rank", str(r),
                               "--rendezvous", str(tmp_path / "nccl_id"), "--bin", str(tmp_path / "state.bin")],
                              cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    shards = []
    for r in range(2):
        raw = np.fromfile(str(tmp_path / "state.bin") + f".rank{r}", dtype="<f8")
        shards.append(raw[: raw.size // 2] + 1j * raw[raw.size // 2:])
    got = np.concatenate(shards)  # canonical layout: rank = top index bit
    fr, fi = G.final_state("mix_n12_f1")
    assert np.max(np.abs(got - (fr + 1j * fi))) < 1e-10
    assert '"exchanges"' in outs[0][0]
