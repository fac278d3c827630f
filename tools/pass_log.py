"""Device time of every pass the library forms from a boundary trace (consecutive dense blocks share a pass).
usage: FLATDD_B200_PASSLOG=1 python tools/pass_log.py <trace-name> [key=value options...]   (lines go to stderr)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import find_trace  # noqa: E402
from flatdd_b200 import Context, read_trace  # noqa: E402

name = sys.argv[1]
opts = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
n, records = read_trace(find_trace(name))
with Context(n) as ctx:
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    gates = [ctx.compile(r.dd) for r in records if r.kind == 2]
    for rep in range(2):
        ctx.set_timing(False)
        ctx.convert(records[0].dd)
        ctx.set_timing(rep == 1)
        ctx.apply_compiled_many(gates)
    print("passes", ctx.get_option("block_launches") // 2, "blocks", ctx.get_option("blocks_applied") // 2, "launches", ctx.launch_count())
