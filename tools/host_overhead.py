"""Host-side cost of fdd_apply per record of a trace: first pass (cold: kernel loading) vs second pass."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context, read_trace
n, records = read_trace(sys.argv[1])
t0 = time.perf_counter()
ctx = Context(n)
print(f"create: {1e3 * (time.perf_counter() - t0):.1f} ms")
for rep in range(2):
    rows = []
    for rec in records:
        t0 = time.perf_counter()
        (ctx.convert if rec.kind == 1 else ctx.apply)(rec.dd)
        ctx.synchronize()
        rows.append(1e3 * (time.perf_counter() - t0))
    print(f"pass {rep}: total {sum(rows):.1f} ms, max {max(rows):.1f} ms, per record:", " ".join(f"{r:.1f}" for r in rows[:14]))
ctx.close()
