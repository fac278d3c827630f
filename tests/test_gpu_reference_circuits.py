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
        re, im = ctx.get_state()
    _check(case, re, im)


@pytest.mark.skipif(not CLI.exists(), reason="build/flatdd_gpu not built")
@pytest.mark.parametrize("fuse", [3, 4])
@pytest.mark.parametrize("case", TRAVEL_CASES)
def test_cli_gpu_fusion(case, fuse):
    m = G.manifest(case, G.TRAVEL)
    circuit = G.circuit_path(m["circuit"])
    if not circuit.exists():
        pytest.skip(f"{circuit} not present")
    out, stats, re, im, _ = run_cli(circuit, 16, fuse, extra=("--quiet",))
    assert stats["switched"] == m["reference"]["switched"]
    _check(case, re, im)
