"""A/B of library builds on the dense upper blocks the tensor-core path serves.
usage: FLATDD_B200_LIB=build/variants/lib_x.so python tools/m5_ab.py <n> [key=value ...]"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context  # noqa: E402
from tests import dd_builder as B  # noqa: E402

n = int(sys.argv[1])
opts = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
shapes = [([10, 11, 12, 13], "dense"), ([7, 12, 17, 22], "dense"), ([n - 1, n - 2, n - 3, n - 4], "dense"), ([10, 11, 12], "dense"),
          ([6, 14, 21], "dense"), ([9, 6, 12, 15, 20], "ctrl")]
rng = np.random.default_rng(0)
out = []
with Context(n) as ctx:
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    yr, yi = B.random_state(n, rng)
    ctx.set_state(yr, yi)
    ctx.set_timing(True)
    for targets, kind in shapes:
        u = B.random_unitary(len(targets), rng) if kind == "dense" else B.controlled(B.random_unitary(len(targets) - 1, rng), 1)
        g = ctx.compile(B.gate_dd(n, targets, u))
        gi = ctx.compile(B.gate_dd(n, targets, u.conj().T))
        times = []
        for _ in range(5):
            ctx.apply_compiled(g)
            times.append(ctx.last_kernel_ms())
            ctx.apply_compiled(gi)
            times.append(ctx.last_kernel_ms())
        ms = float(np.median(times[2:]))
        out.append(f"{str(targets)}/{kind}: {ms:.3f} ms ({32 * (1 << n) / ms / 1e6:.0f} GB/s)")
    probe = ctx.get_amplitudes(12345, 64)
    err = float(np.max(np.abs(probe - (yr[12345:12345 + 64] + 1j * yi[12345:12345 + 64]))))
    tc = ctx.get_option("tensor_core_launches")
print(os.environ.get("FLATDD_B200_LIB", "default"), opts, f"tensor-core launches {tc}, round-trip err {err:.1e}")
for line in out:
    print("   ", line)
