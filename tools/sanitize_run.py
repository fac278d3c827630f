"""Small workload for compute-sanitizer: replays the committed golden traces (n <= 12, every kernel
family incl. forced walk / chunk variants) and checks them against the oracle."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context, read_trace  # noqa: E402
from oracle import pyoracle  # noqa: E402

import os  # noqa: E402

cases = ["tiny_n3_f1", "small_n5_f1", "qft_n8_f1", "mix_n10_f1", "mix_n12_f1", "mix_n12_f0"]
if os.environ.get("SANITIZE_ONLY_NEW"):  # round 1b: only the tensor-core / context-table / flat-table paths below; "2": only the dense-block kernel
    cases = []
for case in cases:
    n, records = read_trace(ROOT / "tests" / "golden" / case / "trace.bin")
    records = records[:40]
    orr, oi = pyoracle.replay_trace(records)
    for variant in (2, 1, 0, 9):
        with Context(n) as ctx:
            ctx.set_option("dmavm_variant", variant)
            for rec in records:
                (ctx.convert if rec.kind == 1 else ctx.apply)(rec.dd)
            re, im = ctx.get_state()
            ctx.norm2()
        err = max(np.max(np.abs(re - orr)), np.max(np.abs(im - oi)))
        assert err < 1e-13, (case, variant, err)
from tests import dd_builder as B  # noqa: E402

# round 2: tile-resident dense-block kernel (warp-specialised: mbarriers, cp.async, setmaxnreg): single blocks on lane / upper / mixed
# targets, controlled blocks, shared passes with and without the CTA barrier between the blocks, a 13-bit tile, a sharded-style rank
rng = np.random.default_rng(1)
block_cases = [
    (14, [([7, 8, 9, 10], 0)], {}), (14, [([0, 1, 2, 3], 0)], {}), (14, [([2, 3, 6, 11], 0)], {"block_tile_bits": 9}), (13, [([4, 5, 6], 0)], {}),
    (14, [([0, 1, 3, 8], 2)], {}), (14, [([9, 10, 13, 6, 7], 3)], {"block_tile_bits": 8}),
    (15, [([5, 6, 7, 8], 0), ([9, 10, 11, 2], 0)], {}), (15, [([0, 1, 6, 7], 0), ([6, 7, 8, 9], 0), ([2, 3, 10], 0)], {}),
    (15, [([0, 1, 2, 3], 0), ([4, 5, 6, 7], 0), ([8, 9, 10, 11], 0)], {}), (15, [([5, 6, 7, 8], 0), ([9, 10, 11, 12], 0)], {"block_max_tile_bits": 13}),
    (15, [([13], 0), ([3, 4], 0), ([5, 6], 0), ([8, 9, 10, 11], 0)], {}),
]
for n, blocks, opts in block_cases:
    gates, ref = [], None
    yr, yi = B.random_state(n, rng)
    psi = yr + 1j * yi
    for qubits, n_ctrl in blocks:
        u = B.random_unitary(len(qubits) - n_ctrl, rng)
        if n_ctrl:
            u = B.controlled(u, n_ctrl)
        gates.append(B.gate_dd(n, qubits, u))
        psi = B.apply_dense(n, qubits, u, psi)
    with Context(n) as ctx:
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.set_state(yr, yi)
        ctx.apply_many(gates)
        re, im = ctx.get_state()
        assert ctx.get_option("blocks_applied") == len(gates), (n, blocks, opts)
    assert np.max(np.abs((re + 1j * im) - psi)) < 1e-13, (n, blocks, opts)
if os.environ.get("SANITIZE_ONLY_NEW") == "2":
    print("sanitize_run ok (dense-block kernel only)")
    sys.exit(0)
# tensor-core path (uniform, context table with 1-3 context bits, per-tile walk), flat-table path, MODE 3

rng = np.random.default_rng(0)
shapes = [(10, [5, 6, 7, 8], 0), (11, [6, 8, 9], 0), (12, [10, 6, 5, 8, 11], 1), (13, [12, 7, 11, 6, 9, 10], 2), (12, [3, 8], 0), (12, [2, 7, 10], 0),
          (12, [0, 3, 9, 11], 0)]
for n, targets, n_ctrl in shapes:
    u = B.random_unitary(len(targets) - n_ctrl, rng)
    if n_ctrl:
        u = B.controlled(u, n_ctrl)
    gate = B.gate_dd(n, targets, u)
    yr, yi = B.random_state(n, rng)
    ref = B.apply_dense(n, targets, u, yr + 1j * yi)
    for opts in ({}, {"context_table": 0}, {"dmma": 0}, {"flat_table": 0}):
        with Context(n) as ctx:
            ctx.set_option("block_kernel", 0)  # these are the round-1 kernels
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.set_state(yr, yi)
            ctx.apply(gate)
            re, im = ctx.get_state()
        assert np.max(np.abs((re + 1j * im) - ref)) < 1e-13, (n, targets, opts)
print("sanitize_run ok")
