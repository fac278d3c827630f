"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C-ABI
(libflatdd_b200.so via ctypes); the oracle (oracle/flat_oracle.c) and the golden vectors written
by the compiled reference are the checkers.

Tolerances (BASELINE.json north_star): state fidelity >= 1 - 1e-10 and max per-amplitude error
<= 1e-10 against the reference.  Conversion is held to bit equality; DMAVM is held to 1e-13
per amplitude here (it differs from the reference only by FMA contraction and the association
order of the path weights)."""
import numpy as np
import pytest

from flatdd_b200 import Context, FlatDDError, load_library, read_flat, read_trace
from oracle import pyoracle
from tests import dd_builder as B
from tests import golden_util as G

pytestmark = pytest.mark.gpu

CASES = G.cases()
AMP_TOL = 1e-13
CONTRACT_AMP_TOL = 1e-10
CONTRACT_INFIDELITY = 1e-10


def _dmavm_kats():
    out = []
    for case in CASES:
        for k in G.manifest(case)["kats"]:
            if k["kind"] == "dmavm":
                out.append((case, k["stem"]))
    return out


def test_library_loads_and_sees_gpu():
    lib = load_library()
    assert lib.device_count() >= 1


@pytest.mark.parametrize("case", CASES)
def test_convert_bit_exact(case):
    """fdd_convert == reference getValueByPathPar bit for bit, every amplitude written."""
    dd = read_flat(G.GOLDEN / case / "kat_convert_dd.bin")
    with Context(dd.n_qubits) as ctx:
        ctx.convert(dd)
        re, im = ctx.get_state()
    assert np.array_equal(re, G.f64(case, "kat_convert_walk_re.f64"))
    assert np.array_equal(im, G.f64(case, "kat_convert_walk_im.f64"))
    # and within rounding of the reference's parallel conversion (regularity shortcut)
    assert G.max_amp_err(re, im, G.f64(case, "kat_convert_switch1_re.f64"), G.f64(case, "kat_convert_switch1_im.f64")) < 1e-15


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("case,stem", _dmavm_kats())
def test_dmavm_kat(case, stem, variant):
    """fdd_apply vs the reference's DDArrMultiplyIP output on the same seeded state."""
    gate = read_flat(G.GOLDEN / case / f"{stem}_dd.bin")
    with Context(gate.n_qubits) as ctx:
        ctx.set_option("dmavm_variant", variant)
        ctx.set_state(G.f64(case, f"{stem}_y_re.f64"), G.f64(case, f"{stem}_y_im.f64"))
        ctx.apply(gate)
        re, im = ctx.get_state()
    err = G.max_amp_err(re, im, G.f64(case, f"{stem}_z_re.f64"), G.f64(case, f"{stem}_z_im.f64"))
    assert err < AMP_TOL, err


@pytest.mark.parametrize("case", CASES)
def test_trace_replay_vs_reference_final_state(case):
    """Whole array phase of a circuit (conversion + every gate the host driver emitted) against
    the reference's own final state."""
    n, records = read_trace(G.GOLDEN / case / "trace.bin")
    with Context(n) as ctx:
        for rec in records:
            if rec.kind == 1:
                ctx.convert(rec.dd)
            else:
                ctx.apply(rec.dd)
        re, im = ctx.get_state()
        norm2 = ctx.norm2()
    fr, fi = G.final_state(case)
    assert G.max_amp_err(re, im, fr, fi) < CONTRACT_AMP_TOL
    assert 1.0 - G.fidelity(re, im, fr, fi) < CONTRACT_INFIDELITY
    assert abs(norm2 - (np.sum(fr * fr) + np.sum(fi * fi))) < 1e-12
    # much tighter in practice: the oracle replay is the same arithmetic up to FMA contraction
    orr, oi = pyoracle.replay_trace(records)
    assert G.max_amp_err(re, im, orr, oi) < 1e-13


@pytest.mark.parametrize("warps,prefetch,ctas", [(1, 1, 1), (2, 2, 0), (4, 4, 2), (8, 8, 0), (16, 16, 0), (8, 16, 1)])
def test_launch_shapes(warps, prefetch, ctas):
    """Result does not depend on the launch configuration."""
    case = "mix_n12_f1"
    n, records = read_trace(G.GOLDEN / case / "trace.bin")
    orr, oi = pyoracle.replay_trace(records)
    with Context(n) as ctx:
        ctx.set_option("warps_per_cta", warps)
        ctx.set_option("prefetch", prefetch)
        ctx.set_option("ctas_per_sm", ctas)
        for rec in records:
            (ctx.convert if rec.kind == 1 else ctx.apply)(rec.dd)
        re, im = ctx.get_state()
    assert G.max_amp_err(re, im, orr, oi) < 1e-13


def test_compiled_gate_replay_and_info():
    case = "mix_n10_f1"
    n, records = read_trace(G.GOLDEN / case / "trace.bin")
    orr, oi = pyoracle.replay_trace(records)
    with Context(n) as ctx:
        gates = [ctx.compile(r.dd) for r in records if r.kind == 2]
        assert all(g.info("max_paths") >= 1 and g.info("max_sub_k") >= 1 for g in gates)
        for _ in range(2):  # the schedule is replayable
            ctx.convert(records[0].dd)
            for g in gates:
                ctx.apply_compiled(g)
            re, im = ctx.get_state()
            assert G.max_amp_err(re, im, orr, oi) < 1e-13
        assert ctx.launch_count() >= 2 * (len(gates) + 1)


def test_ddarr_multiply_dropin():
    """fdd_ddarr_multiply: the literal DDArrMultiplyIP signature on host SoA arrays."""
    case, stem = "mix_n10_f0", "kat_gate1"
    gate = read_flat(G.GOLDEN / case / f"{stem}_dd.bin")
    zr, zi = load_library().ddarr_multiply(gate, G.f64(case, f"{stem}_y_re.f64"), G.f64(case, f"{stem}_y_im.f64"))
    assert G.max_amp_err(zr, zi, G.f64(case, f"{stem}_z_re.f64"), G.f64(case, f"{stem}_z_im.f64")) < AMP_TOL


GATE_SHAPES = [
    # (n, targets, kind)
    (1, [0], "dense"), (2, [1, 0], "dense"), (4, [3], "dense"), (5, [0, 4], "dense"), (6, [5], "dense"),
    (9, [0], "dense"), (12, [4], "dense"), (12, [5], "dense"), (12, [11], "dense"), (13, [2, 3], "dense"),
    (13, [4, 5], "dense"), (13, [0, 12], "dense"), (14, [1, 6, 13], "dense"), (14, [9, 10, 11, 12], "dense"),
    (14, [0, 1, 2, 3, 4], "dense"), (15, [10, 11, 12, 13, 14], "dense"), (15, [14, 0, 7], "ctrl"),
    (15, [0, 14, 7], "ctrl"), (16, [3, 8, 12, 15], "diag"), (16, [15, 2, 9], "perm"), (18, [17, 16, 1, 0], "perm"),
    (20, [19, 5], "dense"), (20, [0, 1, 2, 3, 4, 5], "diag"),
]


@pytest.mark.parametrize("n,targets,kind", GATE_SHAPES)
def test_gate_shapes_vs_numpy_and_oracle(n, targets, kind):
    """Dense / controlled / diagonal / permutation gates on low, high and mixed qubits, checked
    against the oracle and against an independent numpy tensordot."""
    rng = np.random.default_rng(1000 * n + sum(targets) + len(kind))
    k = len(targets)
    if kind == "dense":
        u = B.random_unitary(k, rng)
    elif kind == "ctrl":
        u = B.controlled(B.random_unitary(1, rng), k - 1)
    elif kind == "diag":
        u = np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << k)))
    else:
        u = np.eye(1 << k)[rng.permutation(1 << k)]
    gate = B.gate_dd(n, targets, u)
    yr, yi = B.random_state(n, rng)
    ref = B.apply_dense(n, targets, u, yr + 1j * yi)
    for variant in (0, 1, 2):
        with Context(n) as ctx:
            ctx.set_option("dmavm_variant", variant)
            ctx.set_state(yr, yi)
            ctx.apply(gate)
            re, im = ctx.get_state()
        assert np.max(np.abs((re + 1j * im) - ref)) < AMP_TOL
    if n <= 16:
        orr, oi = pyoracle.dmavm(gate, yr, yi)
        assert G.max_amp_err(re, im, orr, oi) < AMP_TOL


@pytest.mark.parametrize("n,targets", [(12, [11, 10, 9, 8, 7, 6, 5]), (13, [12, 3, 11, 10, 9, 8, 7, 6]), (11, [5, 6, 7, 8, 9, 10])])
def test_very_dense_gate_takes_the_chunk_kernel(n, targets):
    """A block that is dense on many upper qubits (64-256 sources per output segment) exceeds the
    resident lists of the tile and walk kernels; the chunk kernel must still give the exact product."""
    rng = np.random.default_rng(n)
    u = B.random_unitary(len(targets), rng)
    gate = B.gate_dd(n, targets, u)
    yr, yi = B.random_state(n, rng)
    ref = B.apply_dense(n, targets, u, yr + 1j * yi)
    with Context(n) as ctx:
        ctx.set_state(yr, yi)
        ctx.apply(gate)
        re, im = ctx.get_state()
    assert np.max(np.abs((re + 1j * im) - ref)) < 1e-12


def test_512_paths_per_segment():
    """Nine dense upper qubits = 512 source segments per output segment (the reference's own
    --fuse 1 schedule produces such blocks on random circuits): only the chunk kernel holds that."""
    n = 15
    rng = np.random.default_rng(15)
    factors = {q: B.random_unitary(1, rng) for q in range(5, 14)}
    factors[2] = B.random_unitary(1, rng)
    gate = B.kron_dd(n, factors)
    yr, yi = B.random_state(n, rng)
    psi = yr + 1j * yi
    for q, m in factors.items():
        psi = B.apply_dense(n, [q], m, psi)
    with Context(n) as ctx:
        ctx.set_state(yr, yi)
        ctx.apply(gate)
        re, im = ctx.get_state()
    assert np.max(np.abs((re + 1j * im) - psi)) < 1e-12


def test_chunk_kernel_forced_on_a_trace():
    case = "mix_n12_f1"
    n, records = read_trace(G.GOLDEN / case / "trace.bin")
    orr, oi = pyoracle.replay_trace(records)
    with Context(n) as ctx:
        ctx.set_option("dmavm_variant", 9)
        for rec in records:
            (ctx.convert if rec.kind == 1 else ctx.apply)(rec.dd)
        re, im = ctx.get_state()
    assert G.max_amp_err(re, im, orr, oi) < 1e-13


def test_zero_state_and_norm():
    with Context(11) as ctx:
        ctx.set_zero_state()
        re, im = ctx.get_state()
        assert re[0] == 1.0 and np.count_nonzero(re) == 1 and np.count_nonzero(im) == 0
        assert ctx.norm2() == 1.0
        amps = ctx.get_amplitudes(0, 4)
        assert amps[0] == 1.0 and np.all(amps[1:] == 0)


def test_errors_are_loud():
    gate = B.gate_dd(6, [2], B.random_unitary(1, np.random.default_rng(1)))
    with Context(7) as ctx:
        with pytest.raises(FlatDDError):  # no state yet
            ctx.get_state()
        ctx.set_zero_state()
        with pytest.raises(FlatDDError):  # qubit-count mismatch
            ctx.apply(gate)
    bad = B.gate_dd(6, [2], B.random_unitary(1, np.random.default_rng(1)))
    bad.child[bad.root, 0] = 99
    with Context(6) as ctx:
        ctx.set_zero_state()
        with pytest.raises(FlatDDError):
            ctx.apply(bad)


@pytest.mark.parametrize("n", [24, 26])
def test_full_size_properties(n):
    """Size-independent properties at the benchmark's state sizes: unitarity (norm), U^dagger U = 1
    round trip, linearity, and agreement of both kernel variants."""
    rng = np.random.default_rng(n)
    yr, yi = B.random_state(n, rng)
    shapes = [([0], "dense"), ([n - 1], "dense"), ([3, n - 2], "dense"), ([n - 1, 0, 12], "ctrl"), ([1, 7, n - 1], "perm"),
              ([4, 9, 14, n - 3], "diag")]
    with Context(n) as ctx:
        ctx.set_state(yr, yi)
        for targets, kind in shapes:
            k = len(targets)
            if kind == "dense":
                u = B.random_unitary(k, rng)
            elif kind == "ctrl":
                u = B.controlled(B.random_unitary(1, rng), k - 1)
            elif kind == "diag":
                u = np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, size=1 << k)))
            else:
                u = np.eye(1 << k)[rng.permutation(1 << k)]
            ctx.apply(B.gate_dd(n, targets, u))
            assert abs(ctx.norm2() - 1.0) < 1e-12
            # sampled amplitudes against numpy on the touched sub-space is covered at n <= 20;
            # here: undo the gate and compare with the input
            ctx.apply(B.gate_dd(n, targets, u.conj().T))
            idx = rng.integers(0, (1 << n) - 64, size=8)
            for i in idx:
                got = ctx.get_amplitudes(int(i), 64)
                want = yr[i:i + 64] + 1j * yi[i:i + 64]
                assert np.max(np.abs(got - want)) < 1e-13


def test_sampling_follows_the_born_rule():
    """fdd_sample: deterministic in the seed, only populated states, frequencies ~ |amp|^2."""
    n = 14
    rng = np.random.default_rng(5)
    re = np.zeros(1 << n)
    im = np.zeros(1 << n)
    support = rng.choice(1 << n, size=37, replace=False)
    amp = rng.normal(size=37) + 1j * rng.normal(size=37)
    amp /= np.linalg.norm(amp)
    re[support], im[support] = amp.real, amp.imag
    with Context(n) as ctx:
        ctx.set_state(re, im)
        shots = ctx.sample(200000, seed=11)
        again = ctx.sample(200000, seed=11)
        other = ctx.sample(1000, seed=12)
    assert np.array_equal(shots, again) and not np.array_equal(shots[:1000], other)
    assert set(np.unique(shots)) <= set(support.tolist())
    counts = np.array([(shots == s).sum() for s in support]) / shots.size
    assert np.max(np.abs(counts - np.abs(amp) ** 2)) < 0.01
    # a GHZ-like state gives the two extreme outcomes only
    re[:] = 0.0
    im[:] = 0.0
    re[0] = re[-1] = np.sqrt(0.5)
    with Context(n) as ctx:
        ctx.set_state(re, im)
        shots = ctx.sample(4096, seed=1)
    assert set(np.unique(shots)) == {0, (1 << n) - 1}
    assert abs((shots == 0).mean() - 0.5) < 0.05


TENSOR_CORE_SHAPES = [
    # (n, targets, kind): blocks that are complete on 3 or 4 upper qubits with the low five levels untouched
    (9, [5, 6, 7], "dense"), (10, [5, 6, 7, 8], "dense"), (13, [7, 9, 12], "dense"), (14, [13, 9, 6, 5], "dense"),
    (16, [15, 14, 13, 12], "dense"), (17, [5, 8, 11, 16], "dense"),
    (15, [12, 8, 6, 10, 14], "ctrl"),   # control on qubit 12, dense on four others: the block differs from tile to tile
    (15, [13, 7, 9, 11], "ctrl"),       # same with an 8-segment tile
    (14, [6, 8, 10, 12], "half"),       # eight sources per output segment inside a 16-segment tile
    (16, [12, 7, 14, 6, 9, 11], "ctrl3"),  # three controls (one of them a fill bit of the warp tile): 8 context blocks
    (17, [5, 16, 8, 10, 12, 14], "ctrl2"),  # controls on the lowest and the highest segment bit
]


@pytest.mark.parametrize("n,targets,kind", TENSOR_CORE_SHAPES)
def test_tensor_core_path_vs_numpy_and_fma_path(n, targets, kind):
    """Tile kernel MODE 5 (DMMA.8x8x4) against numpy, the oracle and the CUDA-core path of the same gate."""
    rng = np.random.default_rng(77 * n + len(targets))
    k = len(targets)
    if kind == "dense":
        u = B.random_unitary(k, rng)
    elif kind.startswith("ctrl"):
        nc = int(kind[4:] or 1)
        u = B.controlled(B.random_unitary(k - nc, rng), nc)
    else:
        u = np.kron(B.random_unitary(k - 1, rng), np.array([[0.0, 1.0], [1.0, 0.0]]))  # X on the lowest target
    gate = B.gate_dd(n, targets, u)
    yr, yi = B.random_state(n, rng)
    ref = B.apply_dense(n, targets, u, yr + 1j * yi)
    out = {}
    for dmma in (1, 0):
        with Context(n) as ctx:
            ctx.set_option("block_kernel", 0)  # this test is about the tile kernel's paths
            ctx.set_option("dmma", dmma)
            ctx.set_state(yr, yi)
            ctx.apply(gate)
            ctx.apply(gate)  # both ping-pong directions
            out[dmma] = ctx.get_state()
            assert ctx.get_option("tensor_core_launches") == (2 if dmma else 0)
            assert ctx.get_option("context_table_launches") == (2 if dmma else 0)
    if True:  # the walk inside the launch (no context table) gives the same state
        with Context(n) as ctx:
            ctx.set_option("block_kernel", 0)  # this test is about the tile kernel's paths
            ctx.set_option("context_table", 0)
            ctx.set_state(yr, yi)
            ctx.apply(gate)
            ctx.apply(gate)
            walked = ctx.get_state()
            assert ctx.get_option("tensor_core_launches") == 2 and ctx.get_option("context_table_launches") == 0
        assert G.max_amp_err(out[1][0], out[1][1], walked[0], walked[1]) == 0.0
    ref2 = B.apply_dense(n, targets, u, ref)
    for dmma in (1, 0):
        assert np.max(np.abs((out[dmma][0] + 1j * out[dmma][1]) - ref2)) < AMP_TOL
    assert G.max_amp_err(out[1][0], out[1][1], out[0][0], out[0][1]) < 1e-14
    if n <= 14:
        orr, oi = pyoracle.dmavm(gate, yr, yi)
        orr, oi = pyoracle.dmavm(gate, orr, oi)
        assert G.max_amp_err(out[1][0], out[1][1], orr, oi) < AMP_TOL


def test_tensor_core_path_full_size():
    """n = 26 (the benchmark's state): dense 4-qubit block on the tensor cores, undone by its adjoint."""
    n = 26
    rng = np.random.default_rng(2626)
    yr, yi = B.random_state(n, rng)
    targets = [25, 17, 11, 6]
    u = B.random_unitary(4, rng)
    with Context(n) as ctx:
        ctx.set_option("block_kernel", 0)  # this test is about the tile kernel's tensor-core path
        ctx.set_state(yr, yi)
        ctx.apply(B.gate_dd(n, targets, u))
        assert abs(ctx.norm2() - 1.0) < 1e-12
        probe = rng.integers(0, (1 << n) - 64, size=8)
        mid = [ctx.get_amplitudes(int(i), 64) for i in probe]
        ctx.apply(B.gate_dd(n, targets, u.conj().T))
        assert ctx.get_option("tensor_core_launches") == 2
        for i in probe:
            got = ctx.get_amplitudes(int(i), 64)
            assert np.max(np.abs(got - (yr[i:i + 64] + 1j * yi[i:i + 64]))) < 1e-13
        # the same gate on the CUDA-core path gives the same intermediate state
        ctx.set_option("dmma", 0)
        ctx.apply(B.gate_dd(n, targets, u))
        for i, want in zip(probe, mid):
            assert np.max(np.abs(ctx.get_amplitudes(int(i), 64) - want)) < 1e-14


FLAT_TABLE_SHAPES = [
    # (n, targets, expect MODE 6): dense blocks that mix upper and lane qubits, so the sub table depends on the path
    (12, [3, 8], True), (13, [1, 4, 9], True), (14, [2, 7, 11], True), (14, [0, 3, 9, 12], True), (16, [4, 15], True),
    (15, [4, 8, 10, 13], False),  # 8 paths x 2 on an 8-segment tile: table too large, stays on MODE 3
]


@pytest.mark.parametrize("n,targets,flat", FLAT_TABLE_SHAPES)
def test_flat_table_path_vs_numpy_and_mode3(n, targets, flat):
    """Tile kernel MODE 6 (flat precombined table of a uniform gate) against numpy and against MODE 3."""
    rng = np.random.default_rng(31 * n + len(targets))
    u = B.random_unitary(len(targets), rng)
    gate = B.gate_dd(n, targets, u)
    yr, yi = B.random_state(n, rng)
    ref = B.apply_dense(n, targets, u, B.apply_dense(n, targets, u, yr + 1j * yi))
    out = {}
    for on in (1, 0):
        with Context(n) as ctx:
            ctx.set_option("block_kernel", 0)  # this test is about the tile kernel's paths
            ctx.set_option("flat_table", on)
            ctx.set_state(yr, yi)
            ctx.apply(gate)
            ctx.apply(gate)
            out[on] = ctx.get_state()
            assert ctx.get_option("flat_table_launches") == (2 if (on and flat) else 0)
        assert np.max(np.abs((out[on][0] + 1j * out[on][1]) - ref)) < AMP_TOL
    assert G.max_amp_err(out[1][0], out[1][1], out[0][0], out[0][1]) < 1e-14
