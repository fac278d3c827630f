#!/usr/bin/env python
"""bench.py — array-phase throughput of FlatDD's hot path on B200.

Workload (BASELINE.json configs[1]): circuits/supremacy_n26.qasm with DMAVM-aware gate fusion on
one B200.  One STEP = one pass of the hot path over that circuit: the DD->array conversion of the
state DD at the reference's switch point (after op 911) plus one DMAVM launch per fused gate of
the schedule (5078 circuit operations in the array phase).  The inputs are the flat DD tables the
host driver (flatdd_b200/host/gpu_switch_simulator.hpp) emits for that circuit, recorded at build
time into a boundary trace by the product's own binary (`build/flatdd_gpu --trace-only`, committed under
bench_inputs/traces/); nothing here reads /root/reference.

Metric: array-phase circuit operations per second ("gates/s"); seconds per circuit is echoed.
  value  kernels only, gate tables resident in HBM (compiled once), CUDA events on the library stream
  e2e    the same step through the host-buffer C-ABI: fdd_convert + fdd_apply_many per schedule stretch (gate
         compilation and table upload inside) + fdd_get_state into pinned host memory
  roofline  dmavm_tile_kernel: 32 * 2^n algorithmic bytes per launch / mean launch time vs the
         measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference CLI (oracle/_ref/FlatDD) on this box's host cores on a
         bounded sample of the same circuit (first 160 array-phase operations)

`--impl reference` times that reference CLI alone and prints the same line shape.
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOAD = "supremacy_n26"
METRIC = "array_phase_gates_per_sec"
UNIT = "gates/s"


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def find_trace(name: str) -> Path:
    """Boundary traces are made by the product's own host driver (`build/flatdd_gpu --trace-only`,
    tools/make_bench_inputs.py) and committed gzip'd under bench_inputs/traces/."""
    packed = ROOT / "bench_inputs" / "traces" / f"{name}.trace.gz"
    if packed.exists():
        return packed
    raise FileNotFoundError(f"no boundary trace for {name}: run `python tools/make_bench_inputs.py {name}` where build/flatdd_gpu exists")


def table_bytes(dd) -> int:
    return dd.level.nbytes + dd.child.nbytes + dd.weight.nbytes + 32


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi while the timed region runs (B200_PROFILING.md clocks line)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.rows = []
        self._stop = threading.Event()
        self._thread = None
        self._proc = None

    def start(self):
        if shutil.which("nvidia-smi") is None:
            return
        self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

        def pump():
            for line in self._proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self._stop.is_set():
                    break

        self._thread = threading.Thread(target=pump, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=2)
            except Exception:
                self._proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, cell in zip(names, r[5:9]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference CPU arm
# ------------------------------------------------------------------------------------------------
def host_threads() -> int:
    cores = os.cpu_count() or 1
    t = 1
    while t * 2 <= min(cores, 16):
        t *= 2
    return t


def sample_circuit() -> Path:
    p = ROOT / "oracle" / "_ref" / "circuits" / f"{WORKLOAD}_sample.qasm"
    if not p.exists():
        raise FileNotFoundError(f"{p} missing: run `python oracle/make_golden.py samples` where the reference tree exists")
    return p


def run_reference_once(threads: int) -> dict:
    """One run of the unmodified reference CLI on the sample circuit; returns its own timings."""
    exe = ROOT / "oracle" / "_ref" / "FlatDD"
    if not exe.exists():
        raise FileNotFoundError(f"{exe} missing: run `make -C oracle ref` where the reference tree exists")
    circuit = sample_circuit()
    meta = json.loads((circuit.parent / f"{WORKLOAD}_sample.json").read_text())
    with tempfile.TemporaryDirectory() as tmp:
        cwd = Path(tmp) / "build" / "apps"  # the CLI writes to ../../log/results relative to its cwd
        cwd.mkdir(parents=True)
        (Path(tmp) / "log" / "results" / "time").mkdir(parents=True)
        (Path(tmp) / "log" / "results" / "state").mkdir(parents=True)
        t0 = time.perf_counter()
        out = subprocess.run([str(exe), "--file", str(circuit), "-t", str(threads), "--fuse", "1"], cwd=cwd, capture_output=True,
                             text=True, check=True).stdout
        wall = time.perf_counter() - t0
        time_file = next((Path(tmp) / "log" / "results" / "time").glob("*_FlatDD.txt"))
        lines = time_file.read_text().splitlines()
    k = next(i for i, line in enumerate(lines) if line.startswith("Switch Overhead:"))
    convert_s = float(lines[k].split(":")[1])
    array_times = [float(x) for x in lines[k + 1:] if x.strip()]
    switched_at = None
    for line in out.splitlines():
        if line.startswith("Switching at instr."):
            switched_at = int(line.split()[-1])
    array_ops = meta["unitary_ops"] - (switched_at + 1)
    return {"array_s": sum(array_times), "convert_s": convert_s, "launches": len(array_times), "array_ops": array_ops,
            "wall_s": wall, "switched_at": switched_at}


def reference_arm(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    warmup = min(args.warmup, 1)  # a CPU process run has nothing to warm beyond the page cache
    for _ in range(warmup):
        run_reference_once(threads)
    runs = [run_reference_once(threads) for _ in range(args.steps)]
    ops = sum(r["array_ops"] for r in runs)
    secs = sum(r["array_s"] + r["convert_s"] for r in runs)
    value = ops / secs
    sample = (f"first {runs[0]['array_ops']} array-phase ops of {WORKLOAD} after the switch at op {runs[0]['switched_at']} "
              f"({runs[0]['launches']} fused DMAVM calls, conversion included), reference CLI --fuse 1 -t {threads}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warmup, "ms_per_step": 1e3 * secs / max(1, len(runs)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD} array phase (bounded sample)", "n_qubits": 26, "fusion": "reference greedy (--fuse 1)",
                   "host_cores": os.cpu_count(), "threads": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "seconds_per_circuit_extrapolated": 5078 / value,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def trace_name_for(workload: str, world: int) -> str:
    if workload.startswith("supremacy"):  # fused with the GPU cost model
        return f"{workload}_gpu" if world == 1 else f"{workload}_gpu_w{world}"
    if workload.startswith("synth"):
        return f"{workload}_w{world}"
    return f"{workload}_w{world}"  # e.g. knn_n31_f0: per-gate traces exist for every shard count, also _w1


def reference_samples(workload: str):
    """Sampled amplitudes of the reference's own final state (oracle/ref_dump run of the unmodified reference,
    published to bench_inputs/samples/ by `python oracle/make_golden.py publish`): (indices, complex values) or None."""
    import numpy as np
    f = ROOT / "bench_inputs" / "samples" / f"{workload}.samples.bin"
    if not f.exists():
        return None
    raw = f.read_bytes()
    cnt = int(np.frombuffer(raw, dtype="<u8", count=1)[0])
    idx = np.frombuffer(raw, dtype="<u8", count=cnt, offset=8)
    re = np.frombuffer(raw, dtype="<f8", count=cnt, offset=8 + 8 * cnt)
    im = np.frombuffer(raw, dtype="<f8", count=cnt, offset=8 + 16 * cnt)
    return idx, re + 1j * im


def measure(workload: str, args, steps: int, warmup: int, with_e2e: bool) -> dict | None:
    """Times one workload on the ranks of this job; rank 0 gets the result dict, the others None."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from flatdd_b200 import Context, load_library, read_trace

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", local_rank)
    lib = load_library()
    n, records = read_trace(find_trace(trace_name_for(workload, world)))
    assert records[0].kind == 1
    vec = records[0].dd
    array_ops = sum(r.n_original_gates for r in records if r.kind == 2)
    n_local = n - (world.bit_length() - 1)
    local_dim = 1 << n_local

    # N > 1: the state is sharded by its top log2(N) qubits (strong scaling: the same circuit);
    # the trace was scheduled for N shards and carries the half-shard exchanges.
    ctx = Context(n, device=local_rank, rank=rank, world_size=world, library=lib)
    if world > 1:
        uid = [lib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, device=device)
        ctx.comm_init(uid[0])
    for kv in args.option:
        key, value = kv.split("=")
        ctx.set_option(key, int(value))
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
    sched = []  # ("g", [compiled gates of one schedule segment]) | ("x", global bit, local bit) | ("r", a, b)
    for r in records[1:]:
        if r.kind == 2:
            if sched and sched[-1][0] == "g":
                sched[-1][1].append(ctx.compile(r.dd))
            else:
                sched.append(("g", [ctx.compile(r.dd)]))
        elif r.kind == 3:
            sched.append(("x",) + tuple(r.exchange))
        elif r.kind == 4:
            sched.append(("r",) + tuple(r.exchange))
    n_gates = sum(len(s[1]) for s in sched if s[0] == "g")
    n_exch = sum(1 for s in sched if s[0] == "x")
    method = args.exchange_method

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize()

    fuse_exchange = method == 0 and args.fuse_exchange != 0

    def step_resident(exchange_events=None):
        skip = False
        for at, s in enumerate(sched):
            if skip:  # this exchange went along with the stretch before it (fdd_gate_apply_many_exchange)
                skip = False
                continue
            if s[0] == "g":
                if fuse_exchange and at + 1 < len(sched) and sched[at + 1][0] == "x":
                    ctx.apply_compiled_many_exchange(s[1], sched[at + 1][1], sched[at + 1][2])
                    skip = True
                else:
                    ctx.apply_compiled_many(s[1])
            elif s[0] == "x":
                if exchange_events is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    ctx.exchange_qubits(s[1], s[2], method)
                    e1.record(stream)
                    exchange_events.append((e0, e1))
                else:
                    ctx.exchange_qubits(s[1], s[2], method)
            else:
                ctx.relabel_qubits(s[1], s[2])

    for _ in range(warmup):
        ctx.convert(vec)
        step_resident()
    barrier()

    # ---- timed region: K steps, device time on the library's stream ---------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * steps)]
    exchange_events = []
    launches0 = ctx.launch_count()
    tensor0 = ctx.get_option("tensor_core_launches")
    passes0, blocks0 = ctx.get_option("block_launches"), ctx.get_option("blocks_applied")
    fused0 = ctx.get_option("fused_exchanges")
    barrier()
    for s in range(steps):
        ev[3 * s].record(stream)
        ctx.convert(vec)
        ev[3 * s + 1].record(stream)
        step_resident(exchange_events)
        ev[3 * s + 2].record(stream)
    barrier()
    launches = ctx.launch_count() - launches0
    n_fused = (ctx.get_option("fused_exchanges") - fused0) // steps  # exchanges that rode on the pass before them (no kernel of their own)
    # launches that ran FP64 tensor-core code: the older tile kernel's DMMA instantiation and every pass of the dense-block kernel
    tensor_launches = (ctx.get_option("tensor_core_launches") - tensor0) + (ctx.get_option("block_launches") - passes0)
    block_passes = (ctx.get_option("block_launches") - passes0) // steps   # passes of the tile-resident dense-block kernel per step
    blocks_applied = (ctx.get_option("blocks_applied") - blocks0) // steps  # fused gates those passes applied
    # FP64 tensor-core work of a step: a 2^k x 2^k complex block is three real products of 2^k x 2^k x 2^n_local each
    block_flops = 0.0
    for st in sched:
        if st[0] == "g":
            for g in st[1]:
                k = g.info("block_targets")
                if k >= 3:
                    block_flops += 3.0 * 2.0 * (1 << k) * local_dim
    total_ms = ev[0].elapsed_time(ev[3 * steps - 1])
    convert_ms = [ev[3 * s].elapsed_time(ev[3 * s + 1]) for s in range(steps)]
    body_ms = [ev[3 * s + 1].elapsed_time(ev[3 * s + 2]) for s in range(steps)]
    exch_ms = [a.elapsed_time(b) for a, b in exchange_events]
    norm2 = ctx.norm2()

    # ---- e2e: host buffers through the C-ABI, H2D of every table and D2H of the state ------------
    e2e_s, e2e_steps, h2d, d2h, host_norm2 = None, 0, 0, 0, None
    if with_e2e:
        gates = [r.dd for r in records if r.kind == 2]
        h2d = table_bytes(vec) + sum(table_bytes(g) for g in gates)
        download = local_dim if n_local <= 28 else 1 << 20  # very large shards: a 16 MiB sample of the state
        host_re = torch.empty(download, dtype=torch.float64).pin_memory()
        host_im = torch.empty_like(host_re).pin_memory()
        d2h = 16 * download

        def step_e2e():
            # the calls the drop-in driver makes (GpuSwitchSimulator::runSchedule): the host tables of a schedule stretch cross
            # the boundary in one fdd_apply_many call, so its dense blocks share passes exactly as in the device-timed step
            ctx.convert(vec)
            stretch = []
            for r in records[1:]:
                if r.kind == 2:
                    stretch.append(r.dd)
                    continue
                if stretch and r.kind == 3 and fuse_exchange:
                    ctx.apply_many_exchange(stretch, r.exchange[0], r.exchange[1])
                    stretch = []
                    continue
                if stretch:
                    ctx.apply_many(stretch)
                    stretch = []
                if r.kind == 3:
                    ctx.exchange_qubits(r.exchange[0], r.exchange[1], method)
                else:
                    ctx.relabel_qubits(*r.exchange)
            if stretch:
                ctx.apply_many(stretch)
            if world > 1:
                ctx.canonicalize()  # what getVector does for a sharded state: rank r = amplitudes with top index bits r
            if download == local_dim:
                ctx.get_state_raw(host_re.data_ptr(), host_im.data_ptr())
            else:
                amps = ctx.get_amplitudes(0, download)
                host_re.copy_(torch.from_numpy(np.ascontiguousarray(amps.real)))
                host_im.copy_(torch.from_numpy(np.ascontiguousarray(amps.imag)))

        step_e2e()
        barrier()
        e2e_steps = max(1, min(steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        host_norm2 = float(torch.dot(host_re, host_re) + torch.dot(host_im, host_im))
    clocks = sampler.stop()

    # ---- parity inside the bench: the state this run produced against the reference's own sampled amplitudes ----
    ref = reference_samples(workload)
    amp_err, n_checked = None, 0
    if ref is not None:
        if world > 1:
            ctx.canonicalize()
        idx, want = ref
        mine = (idx >> np.uint64(n_local)) == np.uint64(rank)
        got = ctx.get_amplitudes_at(idx[mine] & np.uint64(local_dim - 1))
        amp_err = float(np.max(np.abs(got - want[mine]))) if mine.any() else 0.0
        n_checked = int(mine.sum())

    # ---- max over ranks / sums ---------------------------------------------------------------------
    t_step_ms = total_ms / steps
    exch_mean_ms = statistics.mean(exch_ms) if exch_ms else 0.0
    body_mean = statistics.mean(body_ms)
    if world > 1:
        t = torch.tensor([t_step_ms, e2e_s or 0.0, exch_mean_ms, body_mean, amp_err or 0.0], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step_ms, e2e_max, exch_mean_ms, body_mean, amp_err_max = (float(x) for x in t)
        e2e_s = e2e_max if e2e_s is not None else None
        amp_err = amp_err_max if amp_err is not None else None
        nt = torch.tensor([norm2, host_norm2 or 0.0, float(n_checked)], dtype=torch.float64, device=device)
        dist.all_reduce(nt, op=dist.ReduceOp.SUM)
        norm2, host_sum, n_checked = float(nt[0]), float(nt[1]), int(nt[2])
        host_norm2 = host_sum if host_norm2 is not None else None
    ctx.close()
    if rank != 0:
        return None

    peaks = {}
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peaks = json.loads(peaks_file.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # DMAVM time: the step body minus the exchanges.  One LAUNCH of the dense-block kernel is one pass over the state that
    # applies one or more fused gates to a tile held in shared memory; the algorithmic bytes of a launch are SURVEY.md 8(d)'s
    # per-unit figure (32 * 2^n per applied fused gate) times the fused gates the launch applies.
    dmavm_ms = body_mean - exch_mean_ms * (n_exch - n_fused)  # (a fused exchange is part of its pass: the pass's time includes the transfer)
    launch_ms = dmavm_ms / max(1, n_gates)  # per fused gate
    other_launches = n_gates - blocks_applied  # fused gates that took one of the older kernels: one launch each
    passes = block_passes + other_launches
    achieved = 32.0 * local_dim * n_gates / (dmavm_ms * 1e-3) / 1e9
    dram_gbs = 32.0 * local_dim * passes / (dmavm_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "r02_dmavm_traffic.json"
    if tf.exists() and n_local == 26 and passes > 0:
        table = json.loads(tf.read_text())
        traffic = (table["dmavm_block_ws_kernel"]["dram_bytes_per_launch"] * block_passes +
                   table["dmavm_tile_kernel"]["dram_bytes_per_launch"] * other_launches) / passes
    # TFLOP/s of DMMA.8x8x4 alone on this pool's B200: one DMMA (512 flop) per 16.1-16.4 cycles and scheduler at 1965 MHz
    # (tools/dmma_chain_probe.cu, tools/dadd_probe.cu: 37.0; ncu sm__ops_path_tensor_src_fp64 peak_sustained 128 flop/cycle/SM: 37.2)
    tensor_peak = 37.2
    res = {
        "workload": workload, "n_qubits": n, "n_gates": n_gates, "n_exch": n_exch, "array_ops": array_ops, "n_local": n_local,
        "value": array_ops / (t_step_ms * 1e-3), "ms_per_step": t_step_ms, "convert_ms": statistics.mean(convert_ms),
        "dmavm_ms_per_launch": dmavm_ms / max(1, passes), "dmavm_ms_per_fused_gate": launch_ms, "launches": int(launches), "passes": passes, "tensor_core_launches": int(tensor_launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "dmavm_block_ws_kernel" if block_passes >= other_launches else "dmavm_tile_kernel", "peak_source": peak_src,
                     "launches_per_step": passes, "fused_gates_per_step": n_gates, "fused_gates_per_launch": n_gates / max(1, passes),
                     "bytes_per_launch": 32 * local_dim * n_gates / max(1, passes),
                     "definition": "algorithmic bytes = 32 * 2^n per applied fused gate (SURVEY.md 8d) x the fused gates a launch applies, / launch time; "
                                   "a launch that keeps a tile resident for several fused gates moves 32 * 2^n bytes through HBM ONCE (see dram_*)",
                     "dram_gbs": dram_gbs, "dram_frac": dram_gbs / peak, "dram_bytes_per_launch": 32 * local_dim,
                     "tensor": {"pipe": "FP64 DMMA.8x8x4", "tflops": block_flops / (dmavm_ms * 1e-3) / 1e12, "peak_tflops": tensor_peak,
                                "frac": block_flops / (dmavm_ms * 1e-3) / 1e12 / tensor_peak,
                                "peak_source": "measured: DMMA alone, tools/dmma_chain_probe.cu and tools/dadd_probe.cu (37.0 TFLOP/s = one DMMA per 16.1 cycles and scheduler; ncu's peak_sustained for the pipe: 37.2)"},
                     "convert_gbs": 16.0 * local_dim / (statistics.mean(convert_ms) * 1e-3) / 1e9},
        "check": {"norm2_device": norm2, "max_amp_err_vs_reference": amp_err, "reference_samples_checked": n_checked,
                  "reference": "unmodified reference FlatDD (oracle/ref_dump), sampled amplitudes in bench_inputs/samples/"
                               if amp_err is not None else "no reference samples for this workload (norm only)"},
    }
    if host_norm2 is not None:
        res["check"]["norm2_host_copy"] = host_norm2
    if with_e2e:
        res["e2e"] = {"value": array_ops / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "seconds_per_circuit": e2e_s, "steps": e2e_steps}
    if world > 1 and n_exch:
        half_bytes = 8.0 * local_dim
        gbs = half_bytes / (exch_mean_ms * 1e-3) / 1e9 if exch_mean_ms > 0 else None
        res["exchange"] = {"per_step": n_exch, "fused_into_the_preceding_pass": int(n_fused), "own_kernel": int(n_exch - n_fused),
                           "bytes_each_way_per_gpu": half_bytes, "ms_mean": exch_mean_ms if gbs else None, "gbs_per_direction": gbs,
                           "frac_of_nvlink_nominal_900": gbs / 900.0 if gbs else None, "frac_of_nvlink_measured_770": gbs / 770.0 if gbs else None,
                           "note": "ms / GB/s are those of the exchanges that ran as a kernel of their own; a fused exchange is stored by the pass before it "
                                   "straight into the partner's buffer (peer memory) and has no time of its own",
                           "method": "peer-memory kernels (flags in peer memory, no NCCL call)" if method == 0 else "NCCL send/recv + D2D copy"}
    return res


def circuit_wall(threads: int) -> dict:
    """Whole-circuit wall clock of the drop-in binary: `build/flatdd_gpu --fuse 4` on the .qasm file, the reference's own
    `simulation_time` definition (from before the parse to after simulate(), apps/FlatDD.cpp:46,75,113-120)."""
    exe = ROOT / "build" / "flatdd_gpu"
    circuit = ROOT / "third_party" / "ref_install" / "circuits" / f"{WORKLOAD}.qasm"
    if not exe.exists() or not circuit.exists():
        return {"unavailable": "build/flatdd_gpu or the circuit file is not present (needs `make -C third_party` where a reference checkout exists)"}
    stats = None
    walls = []
    for _ in range(2):  # the first run pages the binary and the CUDA context in
        with tempfile.TemporaryDirectory() as tmp:
            cwd = Path(tmp) / "build" / "apps"
            cwd.mkdir(parents=True)
            (Path(tmp) / "log" / "results" / "time").mkdir(parents=True)
            t0 = time.perf_counter()
            out = subprocess.run([str(exe), "--file", str(circuit), "-t", str(threads), "--fuse", "4", "--quiet"], cwd=cwd,
                                 capture_output=True, text=True, check=True).stdout
            walls.append(time.perf_counter() - t0)
            stats = json.loads(out[out.rindex('{\n  "statistics"'):])["statistics"]
    res = {"binary": "build/flatdd_gpu --fuse 4 (parse + host DD phase + conversion + fusion pass + array phase on the GPU)",
           "simulation_time_s": stats["simulation_time"], "process_wall_s": walls[-1], "gate_merging_s": stats["gate_merging_time"],
           "array_phase_s": stats["array_phase_time"], "dd_to_array_s": stats["DD->Array conversion"],
           "array_phase_launches": stats["array_phase_launches"], "applied_gates": stats["applied_gates"],
           "gates_per_sec_whole_circuit": stats["applied_gates"] / stats["simulation_time"]}
    man = ROOT / "bench_inputs" / "samples" / f"{WORKLOAD}.manifest.json"
    if man.exists():
        ref = json.loads(man.read_text()).get("reference") or {}
        res["reference_simulate_s_recorded"] = ref.get("simulate_s")
        res["reference_note"] = ("the unmodified reference's simulate() on this circuit, --fuse 1 -t 8, recorded when the golden samples were "
                                 "made (not re-run here: 18 minutes)")
    return res


def gpu_arm(args) -> int:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: flatdd_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    workload = args.workload
    warmup = max(3, args.warmup)
    main = measure(workload, args, args.steps, warmup, with_e2e=True)
    # the north-star sharded workloads ride along at N > 1 (fewer steps: their steps are 10-100x longer)
    extras = []
    if args.extra == "auto":
        names = [w for w, counts in (("knn_n31_f0", (2, 4, 8)), ("synth_n34", (8,))) if world in counts] if workload == WORKLOAD else []
    else:
        names = [w for w in args.extra.split(",") if w and w != "none"]
    for name in names:
        extras.append(measure(name, args, steps=3, warmup=3, with_e2e=False))

    if rank == 0:
        n_local = main["n_local"]
        n_exch = main["n_exch"]
        method = args.exchange_method
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{workload} array phase: DD->array conversion + {main['n_gates']} fused gates in {main['passes']} DMAVM launches "
                                   f"({main['array_ops']} circuit ops after the switch)" + (f" + {n_exch} half-shard exchanges" if world > 1 else ""),
                       "n_qubits": main["n_qubits"], "state_bytes": 16 << main["n_qubits"],
                       "fusion": "per gate (fuse 0)" if "_f0" in workload else "dependency-graph fusion with the GPU cost model (fuse 4)",
                       "parallelism": "1 GPU" if world == 1 else f"state sharded over {world} GPUs by its top {world.bit_length() - 1} qubits; "
                                      f"qubit remap + half-shard exchange ({'peer-memory kernel' if method == 0 else 'NCCL send/recv'})",
                       "scaling_note": "the same circuit and state at every GPU count (strong scaling)",
                       "l2": f"shard ({(16 << n_local) >> 20} MiB) and its ping-pong partner exceed the 126 MB L2; no flush needed"
                             if n_local >= 24 else "shard fits L2 at this GPU count (strong scaling of a 1 GiB state)"},
            "seconds_per_circuit": main["ms_per_step"] * 1e-3,
            "convert_ms": main["convert_ms"], "dmavm_ms_per_launch": main["dmavm_ms_per_launch"],
            "roofline": main["roofline"], "e2e": main["e2e"],
            "gpu_launches": main["launches"], "tensor_core_launches": main["tensor_core_launches"],
            "clocks": main["clocks"], "check": main["check"],
        }
        if "exchange" in main:
            line["exchange"] = main["exchange"]
        if extras:
            line["extra_workloads"] = [{k: e[k] for k in ("workload", "n_qubits", "n_gates", "n_exch", "array_ops", "value", "ms_per_step",
                                                           "dmavm_ms_per_launch", "roofline", "check", "clocks") if k in e}
                                       | ({"exchange": e["exchange"]} if "exchange" in e else {}) for e in extras]
        if world == 1 and not args.no_cpu and workload == WORKLOAD:
            try:
                r = run_reference_once(host_threads())
                secs = r["array_s"] + r["convert_s"]
                line["cpu_baseline"] = {
                    "value": r["array_ops"] / secs, "unit": UNIT, "cores": host_threads(), "kind": "reference",
                    "sample": f"first {r['array_ops']} array-phase ops of {WORKLOAD} after the switch at op {r['switched_at']} "
                              f"({r['launches']} fused DMAVM calls + conversion, {secs:.1f} s), unmodified reference CLI --fuse 1 "
                              f"-t {host_threads()} on {os.cpu_count()} host cores"}
            except Exception as exc:  # the baseline is reported, never allowed to void the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "reference", "sample": f"failed: {exc}"}
            try:
                line["circuit_wall"] = circuit_wall(host_threads())
            except Exception as exc:
                line["circuit_wall"] = {"unavailable": f"failed: {exc}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="flatdd_b200", choices=["flatdd_b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default=WORKLOAD, help="supremacy_n26 (default), supremacy_n20, knn_n31_f0, ... (needs its boundary trace)")
    ap.add_argument("--exchange-method", type=int, default=0, help="0 = peer-memory kernel, 1 = NCCL send/recv")
    ap.add_argument("--fuse-exchange", type=int, default=1, help="1: a schedule stretch and the exchange after it cross the boundary in one call (fdd_*_apply_many_exchange)")
    ap.add_argument("--extra", default="auto", help="extra workloads timed in the same run: auto (knn_n31_f0 at 2/4/8 GPUs, synth_n34 at 8), none, or a comma list")
    ap.add_argument("--option", action="append", default=[], help="experiments: library tunable key=value (fdd_set_option), repeatable")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
