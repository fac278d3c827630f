import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context
from tests import dd_builder as B
n = int(sys.argv[1]); targets=[int(x) for x in sys.argv[2].split(",")]
rng = np.random.default_rng(0)
with Context(n) as ctx:
    ctx.set_zero_state()
    g = ctx.compile(B.gate_dd(n, targets, B.random_unitary(len(targets), rng)))
    for _ in range(3): ctx.apply_compiled(g)
    ctx.synchronize()
