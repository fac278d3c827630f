"""GPU parity on the reference's own circuits, against golden data written by the compiled
reference into oracle/_ref/golden/ (git-ignored, travels with the snapshot; regenerate with
`python oracle/make_golden.py medium|large`).  Cases that are not present are skipped.

Two routes per circuit: (a) the boundary trace recorded next to the golden data replayed through
the C-ABI, (b) the flatdd_gpu binary on the .qasm file with the GPU-cost fusions (--fuse 3 and 4), whose
schedule differs from the reference's but whose state must not."""
import numpy as np
import pytest

from flatdd_b200 import Context, read_trace
from tests import golden_util as G
from tests.test_gpu_cli import CLI, ROOT, run_cli

pytestmark = pytest.mark.gpu

TRAVEL_CASES = G.cases(G.TRAVEL)
AMP_TOL = 1e-10
INFIDELITY_TOL = 1e-10


def _check(case, re, im):
    m = G.manifest(case, G.TRAVEL)
    if (G.TRAVEL / case / "final_re.f64").exists():
        fr, fi = G.final_state(case, G.TRAVEL)
        assert G.max_amp_err(re, im, fr, fi) < AMP_TOL
        assert 1.0 - G.fidelity(re, im, fr, fi) < INFIDELITY_TOL
    else:
        idx, sr, si = G.samples(case, G.TRAVEL)
        idx = idx.astype(np.int64)
        assert float(max(np.max(np.abs(re[idx] - sr)), np.max(np.abs(im[idx] - si)))) < AMP_TOL
        norm2 = float(np.dot(re, re) + np.dot(im, im))
        assert abs(norm2 - m["reference"]["norm2"]) < 1e-9


@pytest.mark.parametrize("case", TRAVEL_CASES)
def test_trace_replay(case):
    n, records = read_trace(G.TRAVEL / case / "trace.bin")
    with Context(n) as ctx:
        for rec in records:
            (ctx.convert if rec.kind == 1 else ctx.apply)(rec.dd)
        if n >= 30:
            # knn_n31 (32 GiB): the reference's sampled amplitudes are gathered on the device, the norm is computed there
            m = G.manifest(case, G.TRAVEL)
            idx, sr, si = G.samples(case, G.TRAVEL)
            got = ctx.get_amplitudes_at(idx)
            assert float(np.max(np.abs(got - (sr + 1j * si)))) < AMP_TOL
            # (the reference norm in the manifest is a naive fp64 sum over 2^31 terms in oracle/ref_dump.cpp: good to ~1e-8 only)
            assert abs(ctx.norm2() - m["reference"]["norm2"]) < 1e-8 and abs(ctx.norm2() - 1.0) < 1e-9
            return
        re, im = ctx.get_state()
    _check(case, re, im)


def test_get_amplitudes_at_matches_the_downloaded_state():
    n, records = read_trace(ROOT / "tests" / "golden" / "mix_n10_f1" / "trace.bin")
    with Context(n) as ctx:
        for rec in records:
            (ctx.convert if rec.kind == 1 else ctx.apply)(rec.dd)
        re, im = ctx.get_state()
        idx = np.array([0, 1, 5, 1023, 512, 77, 77], dtype=np.uint64)
        got = ctx.get_amplitudes_at(idx)
        assert np.array_equal(got, re[idx.astype(np.int64)] + 1j * im[idx.astype(np.int64)])
        with pytest.raises(Exception):
            ctx.get_amplitudes_at(np.array([1 << n], dtype=np.uint64))


@pytest.mark.skipif(not CLI.exists(), reason="build/flatdd_gpu not built")
@pytest.mark.parametrize("fuse", [3, 4])
@pytest.mark.parametrize("case", TRAVEL_CASES)
def test_cli_gpu_fusion(case, fuse):
    m = G.manifest(case, G.TRAVEL)
    circuit = G.circuit_path(m["circuit"])
    if not circuit.exists():
        pytest.skip(f"{circuit} not present")
    if m["n_qubits"] >= 30:
        pytest.skip("the CLI route writes the whole state to disk (32 GiB at n = 31); knn_n31 is covered by the trace replay and by bench.py's check")
    out, stats, re, im, _ = run_cli(circuit, 16, fuse, extra=("--quiet",))
    assert stats["switched"] == m["reference"]["switched"]
    _check(case, re, im)
