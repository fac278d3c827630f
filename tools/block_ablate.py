"""Experiments: device time of one dense-block pass (or a shared pass of two) under the kernel's ablation switches.
usage: FLATDD_B200_LIB=build/variants/ablate.so FLATDD_B200_BLOCK_SKIP=<bits> python tools/block_ablate.py <n> <targets;targets...>
bits: 1 no tensor-core work, 2 no global loads/stores, 4 no epilogue (subtractions + stores to the tile), 8 no B-fragment loads"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context  # noqa: E402
from tests import dd_builder as B  # noqa: E402

n = int(sys.argv[1])
groups = [[int(x) for x in g.split(",")] for g in sys.argv[2].split(";")]
rng = np.random.default_rng(0)
with Context(n) as ctx:
    ctx.set_zero_state()
    for kv in os.environ.get("FLATDD_OPTS", "").split(","):
        if kv:
            ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    ctx.set_timing(True)
    gates = [ctx.compile(B.gate_dd(n, t, B.random_unitary(len(t), rng))) for t in groups]
    times = []
    for _ in range(5):
        ctx.apply_compiled_many(gates)
        times.append(ctx.last_kernel_ms())
    print(f"opts={os.environ.get('FLATDD_OPTS', '')} carve={os.environ.get('FLATDD_B200_CARVEOUT', '')} skip={os.environ.get('FLATDD_B200_BLOCK_SKIP', '0'):>2s} lib={Path(os.environ.get('FLATDD_B200_LIB', 'default')).name} targets={sys.argv[2]}: "
          f"{min(times[1:]):.4f} ms (passes {ctx.get_option('block_launches')}, blocks {ctx.get_option('blocks_applied')})", flush=True)
