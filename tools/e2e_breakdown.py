"""Where the end-to-end step (host tables in, state out) spends its time: wall clock of every boundary call with a
synchronize after it.  usage: python tools/e2e_breakdown.py [trace-name]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import find_trace  # noqa: E402
from flatdd_b200 import Context, read_trace  # noqa: E402

n, records = read_trace(find_trace(sys.argv[1] if len(sys.argv) > 1 else "supremacy_n26_gpu"))
gates = [r.dd for r in records if r.kind == 2]
host_re = torch.empty(1 << n, dtype=torch.float64).pin_memory()
host_im = torch.empty_like(host_re).pin_memory()
with Context(n) as ctx:
    for rep in range(3):
        t = [time.perf_counter()]
        ctx.convert(records[0].dd)
        t.append(time.perf_counter())
        ctx.synchronize()
        t.append(time.perf_counter())
        ctx.apply_many(gates)
        t.append(time.perf_counter())
        ctx.synchronize()
        t.append(time.perf_counter())
        ctx.get_state_raw(host_re.data_ptr(), host_im.data_ptr())
        t.append(time.perf_counter())
        ms = [1e3 * (b - a) for a, b in zip(t, t[1:])]
        print(f"rep {rep}: convert call {ms[0]:.2f} + wait {ms[1]:.2f}; apply_many call {ms[2]:.2f} + wait {ms[3]:.2f}; get_state {ms[4]:.2f}; total {sum(ms):.2f} ms")
