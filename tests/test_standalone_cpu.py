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


@pytest.mark.parametrize("fuse", [0, 1, 2])
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
        q.write_text('OPENQASM 2.0;\nqreg q[2];\ngate foo a { h a; }\nfoo q[0];\n')
        with pytest.raises(subprocess.CalledProcessError):
            run_trace_only(q, 0)
