"""Half-shard exchange bandwidth (torchrun, one rank per GPU).
usage: torchrun --nproc-per-node N tools/exchange_bench.py <n_qubits> [key=value ...]"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import Context, load_library  # noqa: E402

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
n = int(sys.argv[1])
opts = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
torch.cuda.set_device(local_rank)
device = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=device)
lib = load_library()
ctx = Context(n, device=local_rank, rank=rank, world_size=world, library=lib)
uid = [lib.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0, device=device)
ctx.comm_init(uid[0])
ctx.set_zero_state()
stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
n_local = ctx.n_local
half_bytes = 8.0 * (1 << n_local)
for key, val in opts.items():
    if key not in ("methods", "reps"):
        ctx.set_option(key, int(val))
methods = [int(m) for m in opts.get("methods", "0,1").split(",")]
reps = int(opts.get("reps", "6"))
for method in methods:
    for pl in sorted({n_local - 1, n_local // 2, 6, 0}):
        for pg in sorted({n_local, n - 1}):
            times = []
            for rep in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ctx.barrier()
                e0.record(stream)
                ctx.exchange_qubits(pg, pl, method)
                e1.record(stream)
                ctx.synchronize()
                times.append(e0.elapsed_time(e1))
            t = torch.tensor([min(times[1:])], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                ms = float(t[0])
                print(f"world={world} n_local={n_local} method={method} global_bit={pg} local_bit={pl}: {ms:.3f} ms  "
                      f"{half_bytes / ms / 1e6:.0f} GB/s per direction ({half_bytes / ms / 1e6 / 900:.2f} of 900)", flush=True)
assert abs(ctx.norm2() - (1.0 if True else 0.0)) < 2.0
ctx.close()
dist.destroy_process_group()
