"""Makes the inputs of bench.py with the product's own host driver: `build/flatdd_gpu --trace-only` runs the host DD
phase, the switch rule and the fusion pass WITHOUT a device and records every flat table that would cross the C-ABI
(with --world N: the N-shard schedule incl. the half-shard exchanges).  The traces are committed gzip'd under
bench_inputs/traces/ so that a box without the reference's front end (parser + DD package, third_party/Makefile) can
still run the bench.

    python tools/make_bench_inputs.py [name ...]        # default: all
"""
from __future__ import annotations

import gzip
import json
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CLI = ROOT / "build" / "flatdd_gpu"
REF_CIRCUITS = ROOT / "third_party" / "ref_install" / "circuits"
OWN_CIRCUITS = ROOT / "bench_inputs" / "circuits"  # tools/gen_synthetic.py
OUT = ROOT / "bench_inputs" / "traces"

# name -> (circuit, fuse, shards)
TRACES = {
    "supremacy_n26_gpu": (REF_CIRCUITS / "supremacy_n26.qasm", 4, 1),
    "supremacy_n26_gpu_w2": (REF_CIRCUITS / "supremacy_n26.qasm", 4, 2),
    "supremacy_n26_gpu_w4": (REF_CIRCUITS / "supremacy_n26.qasm", 4, 4),
    "supremacy_n26_gpu_w8": (REF_CIRCUITS / "supremacy_n26.qasm", 4, 8),
    "knn_n31_f0_w1": (REF_CIRCUITS / "knn_n31.qasm", 0, 1),
    "knn_n31_f0_w2": (REF_CIRCUITS / "knn_n31.qasm", 0, 2),
    "knn_n31_f0_w4": (REF_CIRCUITS / "knn_n31.qasm", 0, 4),
    "knn_n31_f0_w8": (REF_CIRCUITS / "knn_n31.qasm", 0, 8),
    "synth_n30_w8": (OWN_CIRCUITS / "synth_n30.qasm", 4, 8),
    "synth_n34_w8": (OWN_CIRCUITS / "synth_n34.qasm", 4, 8),
}


def make(name: str) -> dict:
    circuit, fuse, world = TRACES[name]
    with tempfile.TemporaryDirectory() as tmp:
        cwd = Path(tmp) / "build" / "apps"  # the CLI writes ../../log/results/... like the reference
        cwd.mkdir(parents=True)
        raw = Path(tmp) / "trace.bin"
        cmd = [str(CLI), "--file", str(circuit), "--fuse", str(fuse), "-t", "8", "--trace", str(raw), "--trace-only", "--quiet"]
        if world > 1:
            cmd += ["--world", str(world)]
        out = subprocess.run(cmd, cwd=cwd, check=True, capture_output=True, text=True).stdout
        meta = json.loads(out[out.index("{"):])["trace"]
        OUT.mkdir(parents=True, exist_ok=True)
        with open(raw, "rb") as src, gzip.GzipFile(OUT / f"{name}.trace.gz", "wb", compresslevel=9, mtime=0) as dst:
            dst.write(src.read())
    meta.pop("file", None)
    meta.pop("simulation_time", None)
    meta.pop("gate_merging_s", None)
    meta["circuit"] = circuit.name
    return meta


def main(argv):
    if not CLI.exists():
        raise SystemExit(f"{CLI} missing: `make -C third_party` where a reference checkout exists")
    names = argv[1:] or list(TRACES)
    index_file = OUT / "index.json"
    index = json.loads(index_file.read_text()) if index_file.exists() else {}
    for name in names:
        index[name] = make(name)
        print(name, index[name], flush=True)
    index_file.write_text(json.dumps(index, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main(sys.argv)
