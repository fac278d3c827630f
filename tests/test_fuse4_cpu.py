"""CPU test of the dependency-graph fusion of the drop-in host driver (GpuSwitchSimulator, --fuse 4): the boundary trace it
emits (DD phase of the reference package, switch, fused gate DDs) replayed on the oracle must reproduce the final state of the
unmodified reference.  Needs oracle/_ref/ref_dump (built where the reference tree exists); skipped elsewhere."""
import subprocess
import tempfile
from pathlib import Path

import pytest

from flatdd_b200 import load_library, read_trace
from oracle import pyoracle
from tests import golden_util as G

ROOT = Path(__file__).resolve().parents[1]
DUMP = ROOT / "oracle" / "_ref" / "ref_dump"

CASES = [("small_n5", "small_n5_f1", 2), ("mix_n7", "mix_n7_f1", 4), ("qft_n8", "qft_n8_f1", 4), ("mix_n10", "mix_n10_f1", 4), ("brick_n11", "brick_n11_f1", 8),
         ("mix_n12", "mix_n12_f1", 8), ("compound_n9", "compound_n9_f1", 4)]


@pytest.mark.skipif(not DUMP.exists(), reason="oracle/_ref/ref_dump not built (needs the reference checkout)")
@pytest.mark.parametrize("name,golden,threads", CASES)
def test_fuse4_trace_reproduces_the_reference_state(name, golden, threads):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([str(DUMP), "--file", str(ROOT / "tests" / "circuits" / f"{name}.qasm"), "--out", tmp, "-t", str(threads), "--fuse", "1",
                        "--trace-fuse", "4", "--no-ref"], check=True, capture_output=True)
        n, records = read_trace(Path(tmp) / "trace.bin")
    m = G.manifest(golden)
    assert n == m["n_qubits"]
    gates = [r for r in records if r.kind == 2]
    # same switch point as the reference's own run (the DD phase and the switch rule do not depend on the fusion mode)
    assert m["reference"]["switched"] == (len(gates) > 0 or m["trace"]["switched_at_op"] >= 0)
    re, im = pyoracle.replay_trace(records)
    fr, fi = G.final_state(golden)
    assert G.max_amp_err(re, im, fr, fi) < 1e-10
    assert 1.0 - G.fidelity(re, im, fr, fi) < 1e-10
    # the policy: at most four non-diagonal qubits above the warp lanes per fused gate
    lib = load_library()
    lanes = min(5, n)
    for r in gates:
        if r.n_original_gates > 1:
            assert bin(lib.matdd_info(r.dd, "non_diag_mask") >> lanes).count("1") <= 4
