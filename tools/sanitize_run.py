"""Small workload for compute-sanitizer: replays the committed golden traces (n <= 12, every kernel
family incl. forced walk / chunk variants) and checks them against the oracle."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context, read_trace  # noqa: E402
from oracle import pyoracle  # noqa: E402

cases = ["tiny_n3_f1", "small_n5_f1", "qft_n8_f1", "mix_n10_f1", "mix_n12_f1", "mix_n12_f0"]
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
print("sanitize_run ok")
