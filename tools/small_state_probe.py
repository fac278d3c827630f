"""Launch cost of the DMAVM kernels on small states (the shards of a 1 GiB state on 8 GPUs hold 2^23 amplitudes and sit in L2).
usage: python tools/small_state_probe.py [n ...]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context  # noqa: E402
from tests import dd_builder as B  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [23, 24, 26]:
    rng = np.random.default_rng(n)
    sets = [sorted(rng.choice(np.arange(5, n), size=4, replace=False).tolist()) for _ in range(12)]
    sets += [sorted(rng.choice(np.arange(0, n), size=4, replace=False).tolist()) for _ in range(12)]
    gates = [B.gate_dd(n, s, B.random_unitary(4, rng)) for s in sets]
    for opts in ({"block_kernel": 1}, {"block_kernel": 1, "block_max_per_pass": 1}, {"block_kernel": 0}):
        with Context(n) as ctx:
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.set_zero_state()
            compiled = [ctx.compile(g) for g in gates]
            stream = torch.cuda.ExternalStream(ctx.stream())
            for _ in range(3):
                ctx.apply_compiled_many(compiled)
            ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = ctx.launch_count()
            e0.record(stream)
            reps = 10
            for _ in range(reps):
                ctx.apply_compiled_many(compiled)
            e1.record(stream)
            ctx.synchronize()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            launches = (ctx.launch_count() - l0) // reps
            print(f"n={n} {opts}: {len(gates)} blocks in {launches} launches, {ms * 1e3 / len(gates):.1f} us per block, "
                  f"{32 * (1 << n) * len(gates) / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic")
