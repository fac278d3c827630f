"""CPU tests of the standalone front end (flatdd_b200/host/standalone.hpp: own OpenQASM reader, gate
matrices, dense-block fusion, flat gate-DD builder).  The binary
runs with --trace-only (no GPU), the oracle replays the boundary trace, and the result is held
against the final state of the UNMODIFIED reference on the same circuit (tests/golden)."""
import json
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from flatdd_b200 import read_trace
from oracle import pyoracle
from tests import golden_util as G

ROOT = Path(__file__).resolve().parents[1]
CLI = ROOT / "build" / "flatdd_gpu_standalone"

CASES = [("tiny_n3", "tiny_n3_f0"), ("small_n5", "small_n5_f0"), ("mix_n7", "mix_n7_f0"), ("qft_n8", "qft_n8_f0"), ("ghz_n6", "ghz_n6_f0"),
         ("mix_n10", "mix_n10_f0"), ("brick_n11", "brick_n11_f1"), ("mix_n12", "mix_n12_f0")]


def build_cli():
    import __graft_entry__ as entry
    entry.build_library()
    return entry.build_standalone()


def run_trace_only(circuit: Path, fuse: int, extra=()):
    with tempfile.TemporaryDirectory() as tmp:
        cwd = Path(tmp) / "build" / "apps"
        cwd.mkdir(parents=True)
        (Path(tmp) / "log" / "results" / "time").mkdir(parents=True)
        trace = Path(tmp) / "trace.bin"
        out = subprocess.run([str(build_cli()), "--file", str(circuit), "--fuse", str(fuse), "--trace", str(trace), "--trace-only", "--quiet", *extra],
                             cwd=cwd, capture_output=True, text=True, check=True).stdout
        n, records = read_trace(trace)
    stats = json.loads(out[out.rindex("\n{\n") + 1:])["statistics"]
    return n, records, stats


@pytest.mark.parametrize("fuse", [0, 1, 2, 3])
@pytest.mark.parametrize("name,golden", CASES)
def test_standalone_trace_reproduces_the_reference_state(name, golden, fuse):
    n, records, stats = run_trace_only(ROOT / "tests" / "circuits" / f"{name}.qasm", fuse)
    m = G.manifest(golden)
    assert n == m["n_qubits"] and stats["applied_gates"] == m["n_ops"]
    assert records[0].kind == 1 and all(r.kind == 2 for r in records[1:])
    assert sum(r.n_original_gates for r in records[1:]) == stats["unitary_gates"]
    if fuse:
        assert len(records) - 1 <= stats["unitary_gates"]
    else:
        assert len(records) - 1 == stats["unitary_gates"]
    re, im = pyoracle.replay_trace(records)
    fr, fi = G.final_state(golden)
    # the reference snaps amplitudes through its DD tolerance before the switch; the contract is 1e-10
    assert G.max_amp_err(re, im, fr, fi) < 1e-10
    assert 1.0 - G.fidelity(re, im, fr, fi) < 1e-10


def test_fusion_policy_bounds_the_blocks():
    """--max-block / --max-nondiag bound every fused block; a Kronecker product of one-qubit gates is one node per level."""
    from flatdd_b200 import load_library
    lib = load_library()
    for fuse in (1, 2):
        for max_block, max_nd in [(2, 2), (4, 3), (5, 4), (6, 4)]:
            n, records, stats = run_trace_only(ROOT / "tests" / "circuits" / "mix_n12.qasm", fuse, ("--max-block", str(max_block), "--max-nondiag", str(max_nd)))
            for r in records[1:]:
                mask = lib.matdd_info(r.dd, "non_diag_mask")
                # a single operation is always allowed (ccx / cswap alone may exceed a tiny policy)
                assert bin(mask).count("1") <= max_nd or r.n_original_gates == 1
    with tempfile.TemporaryDirectory() as tmp:
        q = Path(tmp) / "layer.qasm"
        q.write_text('OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[9];\nry(0.3) q[5];\nrx(0.7) q[6];\nu3(0.1,0.2,0.3) q[7];\nh q[8];\n')
        n, records, _ = run_trace_only(q, 1)
    assert len(records) == 2 and records[1].n_original_gates == 4
    assert records[1].dd.n_nodes == 9  # one node per level: the builder normalises weights towards the root


def test_reader_rejects_what_it_does_not_support():
    with tempfile.TemporaryDirectory() as tmp:
        q = Path(tmp) / "bad.qasm"
        q.write_text('OPENQASM 2.0;\nqreg q[2];\nfoo q[0];\n')
        with pytest.raises(subprocess.CalledProcessError):
            run_trace_only(q, 0)
        q.write_text('OPENQASM 2.0;\nqreg q[2];\ncreg c[2];\nif(c==1) x q[0];\n')
        with pytest.raises(subprocess.CalledProcessError):
            run_trace_only(q, 0)
        q.write_text('OPENQASM 2.0;\nqreg q[2];\ngate foo a { foo a; }\nfoo q[0];\n')  # recursive definition
        with pytest.raises(subprocess.CalledProcessError):
            run_trace_only(q, 0)


def _final_state(qasm_text: str, fuse: int):
    with tempfile.TemporaryDirectory() as tmp:
        q = Path(tmp) / "c.qasm"
        q.write_text(qasm_text)
        n, records, stats = run_trace_only(q, fuse)
    re, im = pyoracle.replay_trace(records)
    return n, re + 1j * im, stats


def test_reader_registers_broadcast_and_expressions():
    """Two quantum registers (first declared = low qubits), whole-register operands, parameter arithmetic."""
    text = ('OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg a[2];\nqreg b[3];\ncreg c[5];\n'
            'h a;\n'                                   # broadcast: h a[0]; h a[1];  (one statement)
            'rx(pi/2) b[0];\nry(-pi/4 + 0.5*2) b[1];\nrz(2^3/(1+3)) b[2];\nu3(sin(0.3),cos(0.2),sqrt(2)) a[1];\n'
            'cx a, b[2];\n'                            # broadcast over the control register
            'barrier a, b;\nmeasure b -> c;\n')
    n, psi, stats = _final_state(text, 0)
    assert n == 5 and stats["applied_gates"] == 10 and stats["unitary_gates"] == 8  # broadcast gates count per element; + barrier + measure
    import math
    from tests import dd_builder as B
    h = np.array([[1, 1], [1, -1]]) / math.sqrt(2)

    def rx(t): return np.array([[math.cos(t / 2), -1j * math.sin(t / 2)], [-1j * math.sin(t / 2), math.cos(t / 2)]])
    def ry(t): return np.array([[math.cos(t / 2), -math.sin(t / 2)], [math.sin(t / 2), math.cos(t / 2)]])
    def rz(t): return np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)])
    def u3(t, p, l): return np.array([[math.cos(t / 2), -np.exp(1j * l) * math.sin(t / 2)], [np.exp(1j * p) * math.sin(t / 2), np.exp(1j * (p + l)) * math.cos(t / 2)]])
    cx = np.eye(4)[[0, 3, 2, 1]]  # dense index: bit 0 = control (targets[0]), bit 1 = target
    ref = np.zeros(32, dtype=complex)
    ref[0] = 1
    for targets, m in [([0], h), ([1], h), ([2], rx(math.pi / 2)), ([3], ry(-math.pi / 4 + 1.0)), ([4], rz(2.0)),
                       ([1], u3(math.sin(0.3), math.cos(0.2), math.sqrt(2))), ([0, 4], cx), ([1, 4], cx)]:
        ref = B.apply_dense(5, targets, m, ref)
    assert np.max(np.abs(psi - ref)) < 1e-14
    for fuse in (1, 2, 3):
        _, fused, _ = _final_state(text, fuse)
        assert np.max(np.abs(fused - ref)) < 1e-14


def test_controlled_and_two_target_gates_match_dense_algebra():
    """ccx / cswap / cu3 / crz / ccz / rzz / rxx / iswap / dcx against an independent numpy construction."""
    import math
    from tests import dd_builder as B
    text = ('OPENQASM 2.0;\nqreg q[6];\n' + "".join(f"ry({0.3 + 0.2 * i}) q[{i}];\n" for i in range(6)) +
            'ccx q[5],q[0],q[3];\ncswap q[1],q[4],q[2];\ncu3(0.4,0.5,0.6) q[3],q[5];\ncrz(0.7) q[0],q[4];\nccz q[2],q[1],q[0];\n'
            'rzz(0.3) q[1],q[5];\nrxx(0.8) q[4],q[0];\niswap q[2],q[3];\ndcx q[5],q[1];\ncy q[0],q[2];\nch q[3],q[1];\nswap q[0],q[5];\n')
    n, psi, _ = _final_state(text, 0)

    def ry(t): return np.array([[math.cos(t / 2), -math.sin(t / 2)], [math.sin(t / 2), math.cos(t / 2)]])
    def u3(t, p, l): return np.array([[math.cos(t / 2), -np.exp(1j * l) * math.sin(t / 2)], [np.exp(1j * p) * math.sin(t / 2), np.exp(1j * (p + l)) * math.cos(t / 2)]])
    x = np.array([[0, 1], [1, 0]], dtype=complex)
    y = np.array([[0, -1j], [1j, 0]])
    z = np.diag([1.0, -1.0]).astype(complex)
    h = np.array([[1, 1], [1, -1]], dtype=complex) / math.sqrt(2)
    rz = lambda t: np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)])
    swap = np.eye(4)[[0, 2, 1, 3]].astype(complex)

    def two(m4):  # reference convention: first listed qubit = high bit; dd_builder: targets[0] = bit 0 -> list the second qubit first
        return m4
    c, s = math.cos(0.15), math.sin(0.15)
    rzz = np.diag([c - 1j * s, c + 1j * s, c + 1j * s, c - 1j * s])
    c, s = math.cos(0.4), math.sin(0.4)
    rxx = np.array([[c, 0, 0, -1j * s], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [-1j * s, 0, 0, c]])
    iswap = np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])
    dcx = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 1, 0, 0], [0, 0, 1, 0]], dtype=complex)
    ref = np.zeros(64, dtype=complex)
    ref[0] = 1
    for i in range(6):
        ref = B.apply_dense(6, [i], ry(0.3 + 0.2 * i), ref)
    steps = [([5, 0, 3], B.controlled(x, 2)), ([1, 2, 4], B.controlled(swap, 1)), ([3, 5], B.controlled(u3(0.4, 0.5, 0.6), 1)),
             ([0, 4], B.controlled(rz(0.7), 1)), ([2, 1, 0], B.controlled(z, 2)),
             ([5, 1], rzz), ([0, 4], rxx), ([3, 2], iswap), ([1, 5], dcx),   # two-target: dense bit 0 = second listed qubit
             ([0, 2], B.controlled(y, 1)), ([3, 1], B.controlled(h, 1)), ([5, 0], swap)]
    for targets, m in steps:
        ref = B.apply_dense(6, targets, m, ref)
    assert np.max(np.abs(psi - ref)) < 1e-14
    for fuse in (1, 2, 3):
        _, fused, _ = _final_state(text, fuse)
        assert np.max(np.abs(fused - ref)) < 1e-13


SHARDED_CASES = [("mix_n10", "mix_n10_f0", 2), ("mix_n10", "mix_n10_f0", 8), ("brick_n11", "brick_n11_f1", 2), ("mix_n12", "mix_n12_f0", 4),
                 ("qft_n8", "qft_n8_f0", 4)]


@pytest.mark.parametrize("name,golden,world", SHARDED_CASES)
def test_standalone_sharded_schedule_on_global_model(name, golden, world):
    """--world N: the standalone front end schedules for N shards (blocks in physical qubit order, Belady remap, half-shard
    exchanges); the trace replayed on the global-array model of tests/test_sharded_cpu.py gives the reference's state."""
    from flatdd_b200.sharded import replay, to_logical_order
    from tests.test_sharded_cpu import GlobalModel
    n, records, stats = run_trace_only(ROOT / "tests" / "circuits" / f"{name}.qasm", 2, ("--world", str(world)))
    n_local = n - int(np.log2(world))
    model = GlobalModel(n, n_local)  # asserts that every gate is diagonal on the global qubits
    l2p = replay(records, model, n)
    got = to_logical_order(model.re + 1j * model.im, l2p)
    fr, fi = G.final_state(golden)
    assert np.max(np.abs(got - (fr + 1j * fi))) < 1e-10
    n_exchanges = sum(r.kind == 3 for r in records)
    assert n_exchanges == stats["exchanges"] and n_exchanges >= 1  # these circuits do act on their top qubits
    for r in records:
        if r.kind == 3:
            assert r.exchange[0] >= n_local > r.exchange[1] >= 0


def test_standalone_sharded_needs_dag_schedule_and_trace_only():
    with tempfile.TemporaryDirectory() as tmp:
        q = ROOT / "tests" / "circuits" / "mix_n10.qasm"
        with pytest.raises(subprocess.CalledProcessError):
            run_trace_only(q, 1, ("--world", "2"))
        cwd = Path(tmp)
        res = subprocess.run([str(build_cli()), "--file", str(q), "--fuse", "2", "--world", "2", "--quiet"], cwd=cwd, capture_output=True, text=True)
        assert res.returncode != 0 and "trace-only" in res.stderr


def test_gate_definitions_expand_like_macros():
    """`gate` blocks: parameters and qubits are substituted at every use, definitions may use each other."""
    import math
    from tests import dd_builder as B
    text = ('OPENQASM 2.0;\ninclude "qelib1.inc";\n'
            'gate rot(theta, phi) a { ry(theta/2) a; rz(phi + pi/4) a; }\n'
            'gate entangle(t) a, b { rot(t, 2*t) a; cx a, b; rot(-t, 0.1) b; barrier a, b; }\n'
            'gate bell a, b { h a; cx a, b; }\n'
            'qreg q[4];\nbell q[0], q[3];\nentangle(0.6) q[1], q[2];\nentangle(pi/3) q[3], q[0];\nrot(0.2, 0.3) q;\n')
    n, psi, stats = _final_state(text, 0)
    assert stats["unitary_gates"] == 2 + 5 + 5 + 8

    def ry(t): return np.array([[math.cos(t / 2), -math.sin(t / 2)], [math.sin(t / 2), math.cos(t / 2)]])
    def rz(t): return np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)])
    h = np.array([[1, 1], [1, -1]]) / math.sqrt(2)
    cx = np.eye(4)[[0, 3, 2, 1]]
    ref = np.zeros(16, dtype=complex)
    ref[0] = 1

    def rot(state, t, p, q):
        state = B.apply_dense(4, [q], ry(t / 2), state)
        return B.apply_dense(4, [q], rz(p + math.pi / 4), state)

    def entangle(state, t, a, b):
        state = rot(state, t, 2 * t, a)
        state = B.apply_dense(4, [a, b], cx, state)
        return rot(state, -t, 0.1, b)
    ref = B.apply_dense(4, [0], h, ref)
    ref = B.apply_dense(4, [0, 3], cx, ref)
    ref = entangle(ref, 0.6, 1, 2)
    ref = entangle(ref, math.pi / 3, 3, 0)
    for q in range(4):
        ref = rot(ref, 0.2, 0.3, q)
    assert np.max(np.abs(psi - ref)) < 1e-14
    _, fused, _ = _final_state(text, 2)
    assert np.max(np.abs(fused - ref)) < 1e-14


@pytest.mark.parametrize("seed", range(8))
def test_fusion_is_state_preserving_on_random_circuits(seed):
    """Random circuits over the whole gate set (1-, 2-, 3-qubit gates, all qubit positions incl. the warp-lane qubits):
    every fusion mode, several policies and a sharded schedule give the state of the unfused run."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_circuits", ROOT / "tests" / "circuits" / "gen_circuits.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    n = 6 + seed % 5  # 6..10 qubits
    text = "\n".join(gen.header(n, f"rand{seed}") + gen.random_circuit(n, 90, seed=1000 + seed) + gen.footer(n)) + "\n"
    _, ref, stats0 = _final_state(text, 0)
    assert abs(np.vdot(ref, ref).real - 1.0) < 1e-12
    with tempfile.TemporaryDirectory() as tmp:
        q = Path(tmp) / "c.qasm"
        q.write_text(text)
        for fuse, extra in [(1, ()), (2, ()), (1, ("--max-block", "3", "--max-nondiag", "2")), (2, ("--max-block", "6", "--max-nondiag", "3")),
                            (2, ("--budget", "1.3")), (3, ())]:
            _, records, stats = run_trace_only(q, fuse, extra)
            re, im = pyoracle.replay_trace(records)
            assert np.max(np.abs((re + 1j * im) - ref)) < 1e-12, (fuse, extra)
            assert stats["array_phase_launches"] <= stats0["array_phase_launches"]
        if n - 1 >= 5:  # two shards need five local qubits
            from flatdd_b200.sharded import replay, to_logical_order
            from tests.test_sharded_cpu import GlobalModel
            for fuse in (2, 3):
                _, records, _ = run_trace_only(q, fuse, ("--world", "2"))
                model = GlobalModel(n, n - 1)
                l2p = replay(records, model, n)
                assert np.max(np.abs(to_logical_order(model.re + 1j * model.im, l2p) - ref)) < 1e-12, fuse
