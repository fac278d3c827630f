"""Runs selected gates of a boundary trace once each (for ncu captures).
usage: python tools/run_gates.py <trace-name> idx [idx...] [key=value options]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import find_trace  # noqa: E402
from flatdd_b200 import Context, read_trace  # noqa: E402

name = sys.argv[1]
idx = [int(a) for a in sys.argv[2:] if "=" not in a]
opts = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
n, records = read_trace(find_trace(name))
gates = [r.dd for r in records if r.kind == 2]
with Context(n) as ctx:
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    ctx.convert(records[0].dd)
    compiled = {i: ctx.compile(gates[i]) for i in idx}
    for rep in range(2):  # first pass warms up, second is the one to look at
        for i in idx:
            ctx.apply_compiled(compiled[i])
    ctx.synchronize()
