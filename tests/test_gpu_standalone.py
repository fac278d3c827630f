"""GPU tests of the standalone front end (build/flatdd_gpu_standalone: own OpenQASM reader, dense-block
fusion, flat start; no reference code at build or run time): whole circuits on the device against the
final states of the unmodified reference."""
import json
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from tests import golden_util as G

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
CLI = ROOT / "build" / "flatdd_gpu_standalone"

SMALL = [("tiny_n3", "tiny_n3_f0"), ("small_n5", "small_n5_f0"), ("mix_n7", "mix_n7_f0"), ("qft_n8", "qft_n8_f0"), ("ghz_n6", "ghz_n6_f0"),
         ("mix_n10", "mix_n10_f0"), ("brick_n11", "brick_n11_f1"), ("mix_n12", "mix_n12_f0")]


def run(circuit: Path, fuse: int, extra=()):
    with tempfile.TemporaryDirectory() as tmp:
        cwd = Path(tmp) / "build" / "apps"
        cwd.mkdir(parents=True)
        (Path(tmp) / "log" / "results" / "time").mkdir(parents=True)
        state = Path(tmp) / "state.bin"
        out = subprocess.run([str(CLI), "--file", str(circuit), "--fuse", str(fuse), "--bin", str(state), "--quiet", *extra],
                             cwd=cwd, capture_output=True, text=True, check=True).stdout
        raw = np.fromfile(state, dtype="<f8")
    full = json.loads(out[out.rindex("\n{\n") + 1:])
    return full, raw[: raw.size // 2], raw[raw.size // 2:]


@pytest.mark.skipif(not CLI.exists(), reason="build/flatdd_gpu_standalone not built")
@pytest.mark.parametrize("fuse", [0, 1, 2, 3])
@pytest.mark.parametrize("name,golden", SMALL)
def test_standalone_final_state(name, golden, fuse):
    full, re, im = run(ROOT / "tests" / "circuits" / f"{name}.qasm", fuse)
    fr, fi = G.final_state(golden)
    assert G.max_amp_err(re, im, fr, fi) < 1e-10
    assert 1.0 - G.fidelity(re, im, fr, fi) < 1e-10
    stats = full["statistics"]
    # (consecutive dense blocks share a pass over the state: fewer kernel launches than fused gates, at least the conversion and one pass)
    assert 2 <= stats["gpu_kernel_launches"] <= stats["array_phase_launches"] + 2 and stats["applied_gates"] == G.manifest(golden)["n_ops"]


TRAVEL = [c for c in G.cases(G.TRAVEL)]


@pytest.mark.skipif(not CLI.exists(), reason="build/flatdd_gpu_standalone not built")
@pytest.mark.parametrize("case", TRAVEL)
def test_standalone_on_reference_circuits(case):
    """The reference's own circuits (golden data written by the compiled reference, oracle/_ref/golden)."""
    m = G.manifest(case, G.TRAVEL)
    circuit = G.circuit_path(m["circuit"])
    if not circuit.exists():
        pytest.skip(f"{circuit} not present")
    if m["n_qubits"] >= 30:
        pytest.skip("this route writes the whole state to disk (32 GiB at n = 31)")
    full, re, im = run(circuit, 2, extra=("--time-gates",))
    if (G.TRAVEL / case / "final_re.f64").exists():
        fr, fi = G.final_state(case, G.TRAVEL)
        assert G.max_amp_err(re, im, fr, fi) < 1e-10
        assert 1.0 - G.fidelity(re, im, fr, fi) < 1e-10
    else:
        idx, sr, si = G.samples(case, G.TRAVEL)
        idx = idx.astype(np.int64)
        assert float(max(np.max(np.abs(re[idx] - sr)), np.max(np.abs(im[idx] - si)))) < 1e-10
        assert abs(float(np.dot(re, re) + np.dot(im, im)) - m["reference"]["norm2"]) < 1e-9
    assert full["statistics"]["dmavm_kernel_ms_total"] > 0


@pytest.mark.skipif(not CLI.exists(), reason="build/flatdd_gpu_standalone not built")
def test_standalone_shots():
    full, re, im = run(ROOT / "tests" / "circuits" / "ghz_n6.qasm", 1, extra=("--shots", "1000", "--seed", "5"))
    assert set(full["samples_top16"].keys()) == {"000000", "111111"} and sum(full["samples_top16"].values()) == 1000


@pytest.mark.skipif(not CLI.exists(), reason="build/flatdd_gpu_standalone not built")
def test_standalone_dump_and_resume():
    """--bin after the first half of a circuit, --load for the second half: the same final state as one run."""
    lines = (ROOT / "tests" / "circuits" / "mix_n10.qasm").read_text().splitlines()
    head = [ln for ln in lines if ln.startswith(("OPENQASM", "include", "qreg", "creg"))]
    body = [ln for ln in lines if ln.strip() and not ln.startswith(("//", "OPENQASM", "include", "qreg", "creg", "measure"))]
    with tempfile.TemporaryDirectory() as tmp:
        first, second = Path(tmp) / "first.qasm", Path(tmp) / "second.qasm"
        first.write_text("\n".join(head + body[: len(body) // 2]) + "\n")
        second.write_text("\n".join(head + body[len(body) // 2:]) + "\n")
        _, re1, im1 = run(first, 2)
        dump = Path(tmp) / "half.bin"
        np.concatenate([re1, im1]).astype("<f8").tofile(dump)
        _, re2, im2 = run(second, 2, extra=("--load", str(dump)))
    _, re, im = run(ROOT / "tests" / "circuits" / "mix_n10.qasm", 2)
    assert G.max_amp_err(re, im, re2, im2) < 1e-13
