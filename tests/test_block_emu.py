"""CPU checks of the tile-resident dense-block path (flatdd_b200/csrc/block_*.{hpp,cpp,cuh}) — no GPU here, so the plan
and the index arithmetic of the kernel are exercised through an emulator (tests/emu/block_emu.cpp) that runs the kernel's
loops lane by lane with the same planner and the same shared helpers, against the oracle's DMAVM."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from flatdd_b200 import read_trace
from flatdd_b200.flat import CMatDD
from oracle import pyoracle
from tests import dd_builder as B
from tests import golden_util as G

ROOT = Path(__file__).resolve().parents[1]
EMU_SRC = ROOT / "tests" / "emu" / "block_emu.cpp"
EMU_SO = ROOT / "tests" / "emu" / "libblock_emu.so"
CSRC = ROOT / "flatdd_b200" / "csrc"


@pytest.fixture(scope="module")
def emu():
    deps = [EMU_SRC, CSRC / "block_compile.cpp", CSRC / "block_compile.hpp", CSRC / "block_plan.hpp", CSRC / "gate_compile.cpp"]
    if not EMU_SO.exists() or any(d.stat().st_mtime > EMU_SO.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", f"-I{ROOT / 'include'}", f"-I{CSRC}", str(EMU_SRC),
                        str(CSRC / "block_compile.cpp"), str(CSRC / "gate_compile.cpp"), "-o", str(EMU_SO)], check=True)
    lib = ctypes.CDLL(str(EMU_SO))
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int)
    lib.emu_apply_pass.argtypes = [ctypes.POINTER(CMatDD), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, ip]
    lib.emu_block_of.argtypes = [ctypes.POINTER(CMatDD), ip, ip, ip, dp, ctypes.c_size_t]
    return lib


def run_pass(lib, gates, n_local, rank, tile_bits, re, im):
    arr = (CMatDD * len(gates))(*[g.as_c() for g in gates])
    info = (ctypes.c_int * 6)()
    re = np.ascontiguousarray(re).copy()
    im = np.ascontiguousarray(im).copy()
    rc = lib.emu_apply_pass(arr, len(gates), n_local, rank, tile_bits, re.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                            im.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), info)
    return rc, re, im, list(info)


def oracle_apply(gates, re, im):
    for g in gates:
        re, im = pyoracle.dmavm(g, re, im)
    return re, im


N = 12
TARGET_SETS = [
    [7, 8, 9, 10],      # upper only: the old tensor-core shape
    [0, 1, 2, 3],       # lane only
    [2, 3, 6, 11],      # lane + upper
    [4, 5, 6],          # k = 3 across the lane boundary
    [0, 11, 5],         # unsorted matrix order
    [1, 9],             # k = 2: padded
    [10],               # k = 1: padded twice
    [0, 3, 6, 9],       # all positions congruent mod 3: the swizzle cannot separate them all
]


@pytest.mark.parametrize("targets", TARGET_SETS)
@pytest.mark.parametrize("tile_bits", [9, 11, 12])
def test_single_block_matches_the_oracle(emu, targets, tile_bits):
    rng = np.random.default_rng(hash(tuple(targets)) % (1 << 31))
    gate = B.gate_dd(N, targets, B.random_unitary(len(targets), rng))
    re, im = B.random_state(N, rng)
    rc, gr, gi, info = run_pass(emu, [gate], N, 0, tile_bits, re, im)
    assert rc == 0
    wr, wi = pyoracle.dmavm(gate, re, im)
    assert max(np.max(np.abs(gr - wr)), np.max(np.abs(gi - wi))) < 1e-14
    assert info[0] <= info[1]  # the planner's conflict estimate is an upper bound of what the emulated accesses show
    if targets != [0, 3, 6, 9]:
        assert info[0] == 1, f"bank conflicts {info}"


@pytest.mark.parametrize("layout", [
    ([3, 8], [0, 1]),        # controls inside the tile (lane bits)
    ([6, 7], [9, 10, 11]),   # controls above: inside or outside the tile depending on its size
    ([2, 9, 10], [4, 11]),
    ([5], [0, 6, 11]),
])
@pytest.mark.parametrize("tile_bits", [8, 10, 12])
def test_controlled_blocks_use_the_context_table(emu, layout, tile_bits):
    targets, controls = layout
    rng = np.random.default_rng(len(targets) * 100 + tile_bits)
    u = B.random_unitary(len(targets), rng)
    # controls are the LOW dense bits of B.controlled
    gate = B.gate_dd(N, controls + targets, B.controlled(u, len(controls)))
    k, n_ctx = block_of(emu, gate)[:2]
    assert k == len(targets) and n_ctx == len(controls)
    re, im = B.random_state(N, rng)
    rc, gr, gi, info = run_pass(emu, [gate], N, 0, tile_bits, re, im)
    assert rc == 0
    wr, wi = pyoracle.dmavm(gate, re, im)
    assert max(np.max(np.abs(gr - wr)), np.max(np.abs(gi - wi))) < 1e-14


def block_of(lib, gate):
    targets = (ctypes.c_int * 8)()
    ctx = (ctypes.c_int * 16)()
    n_ctx = ctypes.c_int(0)
    table = np.zeros(2 * 256 * 1024, dtype=np.float64)
    c = gate.as_c()
    k = lib.emu_block_of(ctypes.byref(c), targets, ctypes.byref(n_ctx), ctx, table.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), table.size)
    return k, n_ctx.value, list(targets)[:max(k, 0)], list(ctx)[:n_ctx.value], table


def test_dense_block_of_a_dd_is_its_matrix(emu):
    """denseBlockFromDD against the dense matrix of the DD itself."""
    rng = np.random.default_rng(5)
    n = 7
    targets, controls = [1, 4, 6], [0, 3]
    u = B.random_unitary(3, rng)
    gate = B.gate_dd(n, controls + targets, B.controlled(u, 2))
    k, n_ctx, tg, cx, table = block_of(emu, gate)
    assert (k, tg, cx) == (3, targets, controls)
    m = table[: 2 * 4 * 8 * 8].reshape(4, 8, 8, 2)
    m = m[..., 0] + 1j * m[..., 1]
    full = gate.to_dense()
    for i in range(1 << n):
        for j in range(1 << n):
            rest_i = i & ~sum(1 << q for q in targets)
            rest_j = j & ~sum(1 << q for q in targets)
            want = 0
            if rest_i == rest_j:
                c = sum(((i >> q) & 1) << a for a, q in enumerate(controls))
                r = sum(((i >> q) & 1) << a for a, q in enumerate(targets))
                cc = sum(((j >> q) & 1) << a for a, q in enumerate(targets))
                want = m[c, r, cc]
            assert abs(full[i, j] - want) < 1e-15


def test_too_wide_gates_are_refused(emu):
    rng = np.random.default_rng(6)
    gate = B.gate_dd(N, [0, 2, 4, 6, 8], B.random_unitary(5, rng))
    assert block_of(emu, gate)[0] == -1


@pytest.mark.parametrize("sets", [
    ([5, 6, 7, 8], [9, 10, 11, 2]),            # a pair on disjoint qubits: 7 upper targets, 12-bit tile
    ([0, 1, 6, 7], [6, 7, 8, 9], [2, 3, 10]),  # overlapping qubits, three blocks
    ([11], [3, 4], [5, 6, 7], [8, 9, 10, 11]),
])
def test_several_blocks_in_one_pass(emu, sets):
    rng = np.random.default_rng(len(sets))
    gates = [B.gate_dd(N, s, B.random_unitary(len(s), rng)) for s in sets]
    re, im = B.random_state(N, rng)
    rc, gr, gi, info = run_pass(emu, gates, N, 0, 12, re, im)
    assert rc == 0
    wr, wi = oracle_apply(gates, re, im)
    assert max(np.max(np.abs(gr - wr)), np.max(np.abs(gi - wi))) < 1e-14
    assert info[5] == 0  # a warp-local pass keeps every compute warp inside its own eighth of the tile


def test_warp_local_passes(emu):
    """Blocks whose targets leave three tile bits free: every block's top unit bits are those bits, so compute warp w only ever
    touches the tile slots with those bits == w and no barrier is needed between the blocks."""
    rng = np.random.default_rng(77)
    sets = ([5, 6, 7, 8], [8, 9, 0, 1], [2, 3, 6])  # union: 0,1,2,3,5,6,7,8,9 -> bits 4, 10, 11 are free
    gates = [B.gate_dd(N, s, B.random_unitary(len(s), rng)) for s in sets]
    re, im = B.random_state(N, rng)
    rc, gr, gi, info = run_pass(emu, gates, N, 0, 12, re, im)
    assert rc == 0 and info[4] == 1 and info[5] == 0
    wr, wi = oracle_apply(gates, re, im)
    assert max(np.max(np.abs(gr - wr)), np.max(np.abs(gi - wi))) < 1e-14
    # every tile bit is some block's target: the pass needs barriers
    full = [B.gate_dd(N, s, B.random_unitary(4, rng)) for s in ([0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11])]
    rc, gr, gi, info = run_pass(emu, full, N, 0, 12, re, im)
    assert rc == 0 and info[4] == 0
    wr, wi = oracle_apply(full, re, im)
    assert max(np.max(np.abs(gr - wr)), np.max(np.abs(gi - wi))) < 1e-14


def test_pass_that_does_not_fit_is_refused(emu):
    rng = np.random.default_rng(9)
    n = 16
    gates = [B.gate_dd(n, s, B.random_unitary(4, rng)) for s in ([5, 6, 7, 8], [9, 10, 11, 12], [13, 14, 15, 0])]
    re, im = B.random_state(n, rng)
    rc, *_ = run_pass(emu, gates, n, 0, 13, re, im)
    assert rc == -2  # 11 upper targets do not fit a 13-bit tile


@pytest.mark.parametrize("world", [2, 4])
def test_shard_with_global_controls(emu, world):
    """Sharded state: the block is controlled by a global qubit (diagonal there); every shard looks its matrix up with its rank."""
    rng = np.random.default_rng(world)
    bits = world.bit_length() - 1
    n_local = N - bits
    targets, controls = [1, 6, n_local - 1], [3, N - 1]
    gate = B.gate_dd(N, controls + targets, B.controlled(B.random_unitary(3, rng), 2))
    re, im = B.random_state(N, rng)
    wr, wi = pyoracle.dmavm(gate, re, im)
    dim = 1 << n_local
    for rank in range(world):
        sl = slice(rank * dim, (rank + 1) * dim)
        rc, gr, gi, _ = run_pass(emu, [gate], n_local, rank, 10, re[sl], im[sl])
        assert rc == 0
        assert max(np.max(np.abs(gr - wr[sl])), np.max(np.abs(gi - wi[sl]))) < 1e-14


@pytest.mark.parametrize("case", ["mix_n10_f1", "mix_n12_f1", "brick_n11_f1", "qft_n8_f1"])
def test_reference_schedules_through_the_emulator(emu, case):
    """The reference's own fused schedules: every gate that is a dense block goes through the emulated kernel, the rest
    through the oracle; the final state is the reference's."""
    n, records = read_trace(G.GOLDEN / case / "trace.bin")
    re, im = pyoracle.convert(records[0].dd)
    as_block = 0
    for rec in records[1:]:
        if n >= 8 and block_of(emu, rec.dd)[0] >= 0:
            rc, re, im, _ = run_pass(emu, [rec.dd], n, 0, min(n, 12), re, im)
            if rc == 0:
                as_block += 1
                continue
        re, im = pyoracle.dmavm(rec.dd, re, im)
    fr, fi = G.final_state(case)
    assert G.max_amp_err(re, im, fr, fi) < 1e-12
    assert as_block > 0
