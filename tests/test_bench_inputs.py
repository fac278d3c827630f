"""The inputs bench.py runs on: committed boundary traces (bench_inputs/traces, made by `build/flatdd_gpu --trace-only`) and the
reference's sampled amplitudes (bench_inputs/samples).  CPU only."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from flatdd_b200 import read_trace
from tests import golden_util as G

ROOT = Path(__file__).resolve().parents[1]
TRACES = ROOT / "bench_inputs" / "traces"
INDEX = json.loads((TRACES / "index.json").read_text())
CLI = ROOT / "build" / "flatdd_gpu"


@pytest.mark.parametrize("name", sorted(INDEX))
def test_trace_matches_its_index(name):
    meta = INDEX[name]
    n, records = read_trace(TRACES / f"{name}.trace.gz")
    assert n == meta["n_qubits"]
    assert records[0].kind == 1 and records[0].dd.radix == 2
    assert sum(r.kind == 2 for r in records) == meta["launches"]
    assert sum(r.kind == 3 for r in records) == meta["exchanges"]
    assert sum(r.n_original_gates for r in records if r.kind == 2) == meta["array_phase_ops"]
    n_local = n - int(np.log2(meta["world"]))
    for r in records:
        if r.kind == 2:  # every gate of a sharded schedule is diagonal on the global qubits
            top = r.dd.level >= n_local
            w = r.dd.weight.reshape(-1, 4, 2)
            assert not w[top][:, 1:3].any()
        if r.kind == 3:
            assert r.exchange[0] >= n_local > r.exchange[1] >= 0


def test_reference_samples_are_the_goldens():
    """bench_inputs/samples/*.samples.bin are byte copies of what oracle/ref_dump wrote next to the reference run."""
    for f in sorted((ROOT / "bench_inputs" / "samples").glob("*.samples.bin")):
        raw = f.read_bytes()
        cnt = int(np.frombuffer(raw, dtype="<u8", count=1)[0])
        assert len(raw) == 8 + 24 * cnt
        man = json.loads(f.with_name(f.name.replace(".samples.bin", ".manifest.json")).read_text())
        assert man["reference"]["switched"] and abs(man["reference"]["norm2"] - 1.0) < 1e-8  # (knn_n31: 4e-9 from the DD tolerance)
        idx = np.frombuffer(raw, dtype="<u8", count=cnt, offset=8)
        assert idx.max() < (1 << man["n_qubits"])


@pytest.mark.skipif(not CLI.exists() or not (G.REF_CIRCUITS / "supremacy_n26.qasm").exists(),
                    reason="needs build/flatdd_gpu and the reference circuits (third_party/Makefile)")
@pytest.mark.parametrize("name,fuse,world", [("supremacy_n26_gpu", 4, 1), ("knn_n31_f0_w8", 0, 8)])
def test_trace_only_cli_reproduces_the_committed_trace(name, fuse, world, tmp_path):
    """No device: the drop-in binary's host side (DD phase, switch rule, fusion pass) writes the same bytes again."""
    import gzip
    cwd = tmp_path / "build" / "apps"
    cwd.mkdir(parents=True)
    circuit = G.REF_CIRCUITS / INDEX[name]["circuit"]
    cmd = [str(CLI), "--file", str(circuit), "--fuse", str(fuse), "-t", "8", "--trace", str(tmp_path / "t.bin"), "--trace-only", "--quiet"]
    if world > 1:
        cmd += ["--world", str(world)]
    out = subprocess.run(cmd, cwd=cwd, check=True, capture_output=True, text=True).stdout
    meta = json.loads(out[out.index("{"):])["trace"]
    assert meta["launches"] == INDEX[name]["launches"] and meta["switched_at_op"] == INDEX[name]["switched_at_op"]
    assert meta["exchanges"] == INDEX[name]["exchanges"]
    if name.startswith("supremacy"):
        assert (tmp_path / "t.bin").read_bytes() == gzip.decompress((TRACES / f"{name}.trace.gz").read_bytes())
    # (knn_n31: the reference's DD package merges nodes within its tolerance in an order that depends on pointer values, so the
    # state DD at the switch has 116 or 142 nodes from run to run; the schedule is the same)
