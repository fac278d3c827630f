"""Times dense k-qubit blocks (random unitaries) on chosen target sets at n qubits.
usage: python tools/dense_bench.py <n> [key=value ...]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context  # noqa: E402
from tests import dd_builder as B  # noqa: E402

n = int(sys.argv[1])
opts = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
shapes = [[10], [2], [10, 11], [3, 10], [2, 3], [10, 11, 12], [3, 10, 11], [2, 3, 10], [10, 11, 12, 13], [3, 10, 11, 12], [2, 3, 10, 11],
          [1, 2, 3, 10], [1, 2, 3, 4], [n - 1, n - 2, n - 3, n - 4], [7, 12, 17, 22], [10, 11, 12, 13, 14], [2, 3, 10, 11, 12], [10, 11, 12, 13, 14, 15]]
rng = np.random.default_rng(0)
with Context(n) as ctx:
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    ctx.set_zero_state()
    ctx.set_timing(True)
    for targets in shapes:
        u = B.random_unitary(len(targets), rng)
        g = ctx.compile(B.gate_dd(n, targets, u))
        times = []
        for _ in range(4):
            ctx.apply_compiled(g)
            times.append(ctx.last_kernel_ms())
        ms = min(times[1:])
        print(f"targets {str(targets):26s} paths={g.info('max_paths'):3d} k={g.info('max_sub_k'):2d} uniform={g.info('uniform')} tb={g.info('sub_tile_bits')}: "
              f"{ms:.3f} ms  {32 * (1 << n) / ms / 1e6:.0f} GB/s", flush=True)
    assert abs(ctx.norm2() - 1.0) < 1e-9
