"""GPU test of the product's host binary (build/flatdd_gpu = reference parser/IR/DD package +
GpuSwitchSimulator + C-ABI library): whole circuits from OpenQASM, final state against the
reference's own final state.  The binary is built where the reference tree exists and travels
with the snapshot; the test is skipped when it is absent."""
import json
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from tests import golden_util as G

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
CLI = ROOT / "build" / "flatdd_gpu"

CASES = [("tiny_n3", 1, 0, "tiny_n3_f0"), ("small_n5", 2, 1, "small_n5_f1"), ("mix_n7", 4, 1, "mix_n7_f1"), ("qft_n8", 4, 0, "qft_n8_f0"),
         ("ghz_n6", 4, 0, "ghz_n6_f0"), ("mix_n10", 4, 0, "mix_n10_f0"), ("mix_n10", 4, 1, "mix_n10_f1"), ("mix_n10", 4, 2, "mix_n10_f2"),
         ("mix_n10", 4, 3, "mix_n10_f1"), ("brick_n11", 8, 3, "brick_n11_f1"), ("mix_n12", 8, 1, "mix_n12_f1"), ("mix_n12", 8, 3, "mix_n12_f1"),
         ("mix_n7", 4, 4, "mix_n7_f1"), ("mix_n10", 4, 4, "mix_n10_f1"), ("brick_n11", 8, 4, "brick_n11_f1"), ("mix_n12", 8, 4, "mix_n12_f1"),
         ("compound_n9", 4, 4, "compound_n9_f1"), ("compound_n9", 4, 0, "compound_n9_f0"),
         # 5: round 1's dependency-graph fusion by DD multiplication
         ("mix_n10", 4, 5, "mix_n10_f1"), ("mix_n12", 8, 5, "mix_n12_f1")]


def run_cli(circuit: Path, threads: int, fuse: int, extra=()):
    with tempfile.TemporaryDirectory() as tmp:
        cwd = Path(tmp) / "build" / "apps"
        cwd.mkdir(parents=True)
        (Path(tmp) / "log" / "results" / "time").mkdir(parents=True)
        (Path(tmp) / "log" / "results" / "state").mkdir(parents=True)
        state = Path(tmp) / "state.bin"
        out = subprocess.run([str(CLI), "--file", str(circuit), "-t", str(threads), "--fuse", str(fuse), "--bin", str(state), *extra],
                             cwd=cwd, capture_output=True, text=True, check=True).stdout
        raw = np.fromfile(state, dtype="<f8")
        time_lines = next((Path(tmp) / "log" / "results" / "time").glob("*_FlatDD.txt")).read_text().splitlines()
    stats = json.loads(out[out.rindex("\n{\n") + 1:])["statistics"]  # the pretty-printed object is the last thing on stdout
    return out, stats, raw[: raw.size // 2], raw[raw.size // 2:], time_lines


@pytest.mark.skipif(not CLI.exists(), reason="build/flatdd_gpu not built (needs the reference checkout)")
@pytest.mark.parametrize("name,threads,fuse,golden", CASES)
def test_cli_final_state_matches_reference(name, threads, fuse, golden):
    out, stats, re, im, time_lines = run_cli(ROOT / "tests" / "circuits" / f"{name}.qasm", threads, fuse)
    fr, fi = G.final_state(golden)
    assert G.max_amp_err(re, im, fr, fi) < 1e-10
    assert 1.0 - G.fidelity(re, im, fr, fi) < 1e-10
    m = G.manifest(golden)
    assert stats["n_qubits"] == m["n_qubits"] and stats["applied_gates"] == m["n_ops"]
    assert stats["switched"] == m["reference"]["switched"]
    if stats["switched"]:
        assert "Switching from DDSIM to FLATDD!!" in out
        if fuse == m["fuse"]:
            assert stats["switched_at_op"] == m["trace"]["switched_at_op"]
    assert any(line.startswith("Switch Overhead:") for line in time_lines)
    assert "Simulation finished" in out


@pytest.mark.skipif(not CLI.exists(), reason="build/flatdd_gpu not built (needs the reference checkout)")
def test_cli_shots_and_time_gates():
    """--shots samples on the device (GHZ: only 000000 and 111111), --time-gates reports kernel time."""
    out, stats, re, im, _ = run_cli(ROOT / "tests" / "circuits" / "ghz_n6.qasm", 4, 0, extra=("--shots", "2000", "--seed", "3"))
    full = json.loads(out[out.rindex("\n{\n") + 1:])
    assert set(full["samples_top16"].keys()) == {"000000", "111111"}
    assert sum(full["samples_top16"].values()) == 2000
    out, stats, re, im, _ = run_cli(ROOT / "tests" / "circuits" / "mix_n12.qasm", 8, 3, extra=("--time-gates", "--quiet"))
    assert stats["dmavm_kernel_ms_total"] > 0 and stats["dmavm_hbm_gbs"] > 0
