"""GPU parity of the tile-resident dense-block kernel (dmavm_block_kernel): through the C-ABI against the oracle, numpy and the
older kernels of the same library.  Tolerance: 1e-13 per amplitude here (contract: 1e-10, fidelity 1 - 1e-10); the block
path differs from the reference by FMA contraction, the three-product complex multiply and the summation order of a row."""
import numpy as np
import pytest

from flatdd_b200 import Context, read_trace
from oracle import pyoracle
from tests import dd_builder as B
from tests import golden_util as G
from tests.test_block_emu import TARGET_SETS

pytestmark = pytest.mark.gpu
AMP_TOL = 1e-13


def apply_all(n, gates, yr, yi, options=(), many=False):
    with Context(n) as ctx:
        for k, v in options:
            ctx.set_option(k, v)
        ctx.set_state(yr, yi)
        if many:
            compiled = [ctx.compile(g) for g in gates]
            ctx.apply_compiled_many(compiled)
        else:
            for g in gates:
                ctx.apply(g)
        re, im = ctx.get_state()
        stats = {k: ctx.get_option(k) for k in ("block_launches", "blocks_applied", "launches")}
    return re, im, stats


@pytest.mark.parametrize("targets", TARGET_SETS)
@pytest.mark.parametrize("tile_bits", [9, 12, 13])
def test_single_block_vs_oracle(targets, tile_bits):
    n = 14
    rng = np.random.default_rng(sum(targets) * 7 + tile_bits)
    gate = B.gate_dd(n, targets, B.random_unitary(len(targets), rng))
    yr, yi = B.random_state(n, rng)
    re, im, stats = apply_all(n, [gate, gate], yr, yi, [("block_tile_bits", tile_bits)])  # twice: both ping-pong directions
    assert stats["block_launches"] == 2
    wr, wi = pyoracle.dmavm(gate, *pyoracle.dmavm(gate, yr, yi))
    assert G.max_amp_err(re, im, wr, wi) < AMP_TOL


@pytest.mark.parametrize("layout", [([3, 8], [0, 1]), ([6, 7], [9, 10, 13]), ([2, 9, 10], [4, 11]), ([5], [0, 6, 12]), ([1, 2, 3, 4], [5, 7, 9, 11, 13])])
@pytest.mark.parametrize("tile_bits", [8, 11, 13])
def test_controlled_block_vs_oracle_and_old_kernels(layout, tile_bits):
    n = 14
    targets, controls = layout
    rng = np.random.default_rng(len(controls) * 50 + tile_bits)
    gate = B.gate_dd(n, controls + targets, B.controlled(B.random_unitary(len(targets), rng), len(controls)))
    yr, yi = B.random_state(n, rng)
    re, im, stats = apply_all(n, [gate], yr, yi, [("block_tile_bits", tile_bits)])
    assert stats["block_launches"] == 1
    wr, wi = pyoracle.dmavm(gate, yr, yi)
    assert G.max_amp_err(re, im, wr, wi) < AMP_TOL
    ore, oim, old = apply_all(n, [gate], yr, yi, [("block_kernel", 0)])
    assert old["block_launches"] == 0
    assert G.max_amp_err(re, im, ore, oim) < AMP_TOL


@pytest.mark.parametrize("sets,passes", [
    (([5, 6, 7, 8], [9, 10, 11, 2]), 1),                 # 7 upper targets: one 12-bit tile
    (([5, 6, 7, 8], [9, 10, 11, 12]), 2),                # 8 upper targets would need a 13-bit tile (one buffer): two passes by default
    (([5, 6, 7, 8], [9, 10, 11, 2], [13, 0, 1]), 2),     # the third block does not fit the tile of the first two
    (([0, 1, 6, 7], [6, 7, 8, 9], [2, 3, 10]), 1),
    (([13], [3, 4], [5, 6], [8, 9, 10, 11]), 1),         # seven upper targets between four blocks (the small ones are padded with lane qubits)
])
def test_blocks_share_a_pass(sets, passes):
    n = 15
    rng = np.random.default_rng(len(sets) + passes)
    gates = [B.gate_dd(n, s, B.random_unitary(len(s), rng)) for s in sets]
    yr, yi = B.random_state(n, rng)
    re, im, stats = apply_all(n, gates, yr, yi, many=True)
    assert stats["blocks_applied"] == len(sets) and stats["block_launches"] == passes
    wr, wi = yr, yi
    for g in gates:
        wr, wi = pyoracle.dmavm(g, wr, wi)
    assert G.max_amp_err(re, im, wr, wi) < AMP_TOL
    # one pass per block gives the same state up to rounding
    sre, sim, single = apply_all(n, gates, yr, yi, [("block_max_per_pass", 1)], many=True)
    assert single["block_launches"] == len(sets)
    assert G.max_amp_err(re, im, sre, sim) < 1e-14


def test_commuting_blocks_move_up_to_share_a_pass():
    """orderForPasses (api.cu): A, B, A', B' with A and B on disjoint upper qubits (eight between them: no common tile) run as the
    passes [A, A'] and [B, B'] because A' commutes with B; without the reordering every block has its own pass."""
    n = 15
    rng = np.random.default_rng(11)
    sets = ([5, 6, 7, 8], [9, 10, 11, 12], [8, 7, 6, 5], [12, 9, 10, 11])
    gates = [B.gate_dd(n, s, B.random_unitary(len(s), rng)) for s in sets]
    yr, yi = B.random_state(n, rng)
    wr, wi = yr, yi
    for g in gates:
        wr, wi = pyoracle.dmavm(g, wr, wi)
    re, im, stats = apply_all(n, gates, yr, yi, many=True)
    assert stats["blocks_applied"] == 4 and stats["block_launches"] == 2
    assert G.max_amp_err(re, im, wr, wi) < AMP_TOL
    ore, oim, plain = apply_all(n, gates, yr, yi, [("block_reorder", 0)], many=True)
    assert plain["block_launches"] == 4
    assert G.max_amp_err(ore, oim, wr, wi) < AMP_TOL


def test_blocks_that_do_not_commute_keep_their_order():
    """C shares qubit 12 with B, so it may not jump over B to join A's pass; a control qubit shared by two gates does not stop them
    (both are diagonal in it), a control that is another gate's target does."""
    n = 15
    rng = np.random.default_rng(12)
    a = B.gate_dd(n, [5, 6, 7, 8], B.random_unitary(4, rng))
    b = B.gate_dd(n, [9, 10, 11, 12], B.random_unitary(4, rng))
    c = B.gate_dd(n, [5, 6, 7, 12], B.random_unitary(4, rng))
    d = B.gate_dd(n, [3, 13, 14], B.controlled(B.random_unitary(2, rng), 1))   # control 3
    e = B.gate_dd(n, [3, 5, 6], B.controlled(B.random_unitary(2, rng), 1))     # control 3 as well: commutes with d
    f = B.gate_dd(n, [3, 4], B.random_unitary(2, rng))                         # target 3: commutes with neither
    for gates in ([a, b, c], [d, b, e], [d, f, e], [a, d, b, f, c, e]):
        yr, yi = B.random_state(n, rng)
        wr, wi = yr, yi
        for g in gates:
            wr, wi = pyoracle.dmavm(g, wr, wi)
        re, im, _ = apply_all(n, gates, yr, yi, many=True)
        assert G.max_amp_err(re, im, wr, wi) < AMP_TOL


@pytest.mark.parametrize("seed", range(6))
def test_random_block_sequences_reordered_vs_oracle(seed):
    n = 15
    rng = np.random.default_rng(100 + seed)
    gates = []
    for _ in range(12):
        k = int(rng.integers(1, 5))
        qubits = [int(q) for q in rng.choice(n, size=k + int(rng.integers(0, 3)), replace=False)]
        controls, targets = qubits[k:], qubits[:k]
        u = B.random_unitary(k, rng)
        gates.append(B.gate_dd(n, controls + targets, B.controlled(u, len(controls)) if controls else u))
    yr, yi = B.random_state(n, rng)
    wr, wi = yr, yi
    for g in gates:
        wr, wi = pyoracle.dmavm(g, wr, wi)
    re, im, stats = apply_all(n, gates, yr, yi, many=True)
    assert G.max_amp_err(re, im, wr, wi) < AMP_TOL
    ore, oim, plain = apply_all(n, gates, yr, yi, [("block_reorder", 0)], many=True)
    assert G.max_amp_err(ore, oim, wr, wi) < AMP_TOL
    assert stats["block_launches"] <= plain["block_launches"]
    # the matrix tables of a shared pass from global memory instead of shared memory: the same arithmetic
    gre, gim, _ = apply_all(n, gates, yr, yi, [("block_tables_shared", 0)], many=True)
    assert G.max_amp_err(gre, gim, re, im) == 0.0


def test_wide_gate_falls_back_to_the_older_kernels():
    n = 14
    rng = np.random.default_rng(3)
    gate = B.gate_dd(n, [0, 3, 6, 9, 12], B.random_unitary(5, rng))
    yr, yi = B.random_state(n, rng)
    re, im, stats = apply_all(n, [gate], yr, yi)
    assert stats["block_launches"] == 0 and stats["launches"] >= 2
    wr, wi = pyoracle.dmavm(gate, yr, yi)
    assert G.max_amp_err(re, im, wr, wi) < AMP_TOL


@pytest.mark.parametrize("case", [c for c in G.cases() if G.manifest(c)["n_qubits"] >= 8])
def test_reference_schedules_on_the_block_path(case):
    """Traces of the reference's own schedules: same final state with and without the block kernel, both the reference's."""
    n, records = read_trace(G.GOLDEN / case / "trace.bin")
    out = {}
    for on in (1, 0):
        with Context(n) as ctx:
            ctx.set_option("block_kernel", on)
            for rec in records:
                (ctx.convert if rec.kind == 1 else ctx.apply)(rec.dd)
            out[on] = ctx.get_state()
            blocks = ctx.get_option("block_launches")
        assert (blocks > 0) == bool(on) or not any(r.kind == 2 for r in records)
    fr, fi = G.final_state(case)
    assert G.max_amp_err(out[1][0], out[1][1], fr, fi) < 1e-12
    assert G.max_amp_err(out[1][0], out[1][1], out[0][0], out[0][1]) < 1e-13


def test_thirteen_bit_tile_when_asked_for():
    """block_max_tile_bits = 13: eight upper targets share one pass (single tile buffer)."""
    n = 15
    rng = np.random.default_rng(13)
    sets = ([5, 6, 7, 8], [9, 10, 11, 12])
    gates = [B.gate_dd(n, s, B.random_unitary(len(s), rng)) for s in sets]
    yr, yi = B.random_state(n, rng)
    re, im, stats = apply_all(n, gates, yr, yi, [("block_max_tile_bits", 13)], many=True)
    assert stats["block_launches"] == 1 and stats["blocks_applied"] == 2
    wr, wi = yr, yi
    for g in gates:
        wr, wi = pyoracle.dmavm(g, wr, wi)
    assert G.max_amp_err(re, im, wr, wi) < AMP_TOL


def test_full_size_pair_round_trip():
    """n = 26: two dense 4-qubit blocks with seven upper qubits between them in one pass, undone by their adjoints."""
    n = 26
    rng = np.random.default_rng(2627)
    yr, yi = B.random_state(n, rng)
    ta, tb = [25, 17, 11, 6], [22, 9, 14, 3]
    ua, ub = B.random_unitary(4, rng), B.random_unitary(4, rng)
    with Context(n) as ctx:
        ctx.set_state(yr, yi)
        fwd = [ctx.compile(B.gate_dd(n, ta, ua)), ctx.compile(B.gate_dd(n, tb, ub))]
        bwd = [ctx.compile(B.gate_dd(n, tb, ub.conj().T)), ctx.compile(B.gate_dd(n, ta, ua.conj().T))]
        ctx.apply_compiled_many(fwd)
        assert abs(ctx.norm2() - 1.0) < 1e-12
        probe = rng.integers(0, (1 << n) - 64, size=8)
        mid = [ctx.get_amplitudes(int(i), 64) for i in probe]
        ctx.apply_compiled_many(bwd)
        assert ctx.get_option("block_launches") == 2 and ctx.get_option("blocks_applied") == 4
        for i in probe:
            assert np.max(np.abs(ctx.get_amplitudes(int(i), 64) - (yr[i:i + 64] + 1j * yi[i:i + 64]))) < 1e-13
        # the same two gates through the older kernels give the same intermediate state
        ctx.set_option("block_kernel", 0)
        ctx.apply_compiled_many(fwd)
        for i, want in zip(probe, mid):
            assert np.max(np.abs(ctx.get_amplitudes(int(i), 64) - want)) < 1e-13


def test_lane_qubit_block_full_size():
    """n = 26: a block that mixes warp-lane qubits with upper qubits (the old kernels' slow classes) on the tensor cores."""
    n = 26
    rng = np.random.default_rng(2628)
    yr, yi = B.random_state(n, rng)
    targets = [1, 3, 12, 20]
    u = B.random_unitary(4, rng)
    with Context(n) as ctx:
        ctx.set_state(yr, yi)
        ctx.apply(B.gate_dd(n, targets, u))
        assert ctx.get_option("block_launches") == 1
        assert abs(ctx.norm2() - 1.0) < 1e-12
        probe = rng.integers(0, (1 << n) - 64, size=8)
        got = [ctx.get_amplitudes(int(i), 64) for i in probe]
        ctx.set_state(yr, yi)
        ctx.set_option("block_kernel", 0)
        ctx.apply(B.gate_dd(n, targets, u))
        for i, want in zip(probe, got):
            assert np.max(np.abs(ctx.get_amplitudes(int(i), 64) - want)) < 1e-13
