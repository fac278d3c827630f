"""Replays a trace gate by gate on the GPU and on the CPU oracle; prints the first gate whose
result deviates, with its structural facts.  usage: python tools/debug_trace.py <trace.bin> [key=value]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context, load_library, read_trace  # noqa: E402
from oracle import pyoracle  # noqa: E402

path = sys.argv[1]
opts = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
n, records = read_trace(path)
lib = load_library()
keys = ["max_paths", "max_sub_k", "upper_depth", "upper_nodes", "sub_tables", "tileable", "sub_tile_bits", "non_diag_upper", "stack_cap"]
with Context(n) as ctx:
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    re = im = None
    bad = 0
    for idx, rec in enumerate(records):
        if rec.kind == 1:
            ctx.convert(rec.dd)
            re, im = pyoracle.convert(rec.dd)
            continue
        ctx.apply(rec.dd)
        re, im = pyoracle.dmavm(rec.dd, re, im)
        gr, gi = ctx.get_state()
        err = max(np.max(np.abs(gr - re)), np.max(np.abs(gi - im)))
        if err > 1e-12:
            print(f"record {idx}: err {err:.3e}", {k: lib.matdd_info(rec.dd, k) for k in keys}, "nodes", rec.dd.n_nodes, flush=True)
            bad += 1
            ctx.set_state(re, im)  # resync so later gates are judged on their own
            if bad >= 4:
                break
    print("done, bad gates:", bad)
