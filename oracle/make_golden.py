"""Regenerates the golden fixtures by running the compiled reference (oracle/_ref/ref_dump).

    python oracle/make_golden.py small     # tests/golden/<case>/   (committed)
    python oracle/make_golden.py medium    # oracle/_ref/golden/<case>/ (git-ignored, travels with gpurun)
    python oracle/make_golden.py large     # samples only, minutes to tens of minutes of CPU
    python oracle/make_golden.py sharded   # tests/golden_sharded/<case>/: boundary traces scheduled for 2/4/8 shards
    python oracle/make_golden.py traces    # boundary traces of the bench circuits (no reference run)
    python oracle/make_golden.py samples   # truncated circuits for the bounded CPU-baseline runs
    python oracle/make_golden.py circuits  # copy the reference's 12 circuits next to the binaries (git-ignored)
    python oracle/make_golden.py publish   # copy the reference's sampled amplitudes of the bench workloads to bench_inputs/samples/ (committed)

Needs /root/reference (for `make -C oracle ref`) and the repo's own test circuits."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
DUMP = ROOT / "oracle" / "_ref" / "ref_dump"

# case name -> (circuit, threads, fuse, extra args)
SMALL = {
    "tiny_n3_f0": ("tests/circuits/tiny_n3.qasm", 1, 0, ["--kat-gates", "3"]),
    "tiny_n3_f1": ("tests/circuits/tiny_n3.qasm", 1, 1, ["--kat-gates", "2"]),
    "small_n5_f0": ("tests/circuits/small_n5.qasm", 2, 0, ["--kat-gates", "3"]),
    "small_n5_f1": ("tests/circuits/small_n5.qasm", 2, 1, ["--kat-gates", "2"]),
    "mix_n7_f0": ("tests/circuits/mix_n7.qasm", 4, 0, ["--kat-gates", "4"]),
    "mix_n7_f1": ("tests/circuits/mix_n7.qasm", 4, 1, ["--kat-gates", "2"]),
    "qft_n8_f0": ("tests/circuits/qft_n8.qasm", 4, 0, ["--kat-gates", "3"]),
    "qft_n8_f1": ("tests/circuits/qft_n8.qasm", 4, 1, ["--kat-gates", "2"]),
    "ghz_n6_f0": ("tests/circuits/ghz_n6.qasm", 4, 0, ["--kat-gates", "2"]),
    "mix_n10_f0": ("tests/circuits/mix_n10.qasm", 4, 0, ["--kat-gates", "4"]),
    "mix_n10_f1": ("tests/circuits/mix_n10.qasm", 4, 1, ["--kat-gates", "2"]),
    "mix_n10_f1nc": ("tests/circuits/mix_n10.qasm", 4, 1, ["--kat-gates", "1", "--no_cache"]),
    "mix_n10_f2": ("tests/circuits/mix_n10.qasm", 4, 2, ["--kat-gates", "1"]),
    "brick_n11_f1": ("tests/circuits/brick_n11.qasm", 8, 1, ["--kat-gates", "1"]),
    "mix_n12_f0": ("tests/circuits/mix_n12.qasm", 8, 0, ["--kat-gates", "1"]),
    "mix_n12_f1": ("tests/circuits/mix_n12.qasm", 8, 1, ["--kat-gates", "1"]),
    # user-defined gates (compound operations) with SWAPs inside
    "compound_n9_f0": ("tests/circuits/compound_n9.qasm", 4, 0, ["--kat-gates", "1"]),
    "compound_n9_f1": ("tests/circuits/compound_n9.qasm", 4, 1, ["--kat-gates", "1"]),
}
MEDIUM = {
    "synth_n20_f1": (ROOT / "bench_inputs/circuits/synth_n20.qasm", 8, 1, ["--no-kat", "--samples", "65536"]),
    "vqe_n16_f0": (REF / "circuits/vqe_n16.qasm", 8, 0, ["--no-kat", "--full-state"]),
    "vqe_n16_f1": (REF / "circuits/vqe_n16.qasm", 8, 1, ["--no-kat", "--full-state"]),
    "dnn_n16_f1": (REF / "circuits/dnn_n16.qasm", 8, 1, ["--no-kat", "--full-state"]),
    "ghz_state_n23_f0": (REF / "circuits/ghz_state_n23.qasm", 8, 0, ["--no-kat"]),
    "supremacy_n20_f1": (REF / "circuits/supremacy_n20.qasm", 8, 1, ["--no-kat", "--samples", "65536"]),
    "dnn_n20_f1": (REF / "circuits/dnn_n20.qasm", 8, 1, ["--no-kat"]),
}
LARGE = {
    "adder_n28_f0": (REF / "circuits/adder_n28.qasm", 8, 0, ["--no-kat"]),
    "knn_n25_f1": (REF / "circuits/knn_n25.qasm", 8, 1, ["--no-kat"]),
    "swap_test_n25_f1": (REF / "circuits/swap_test_n25.qasm", 8, 1, ["--no-kat"]),
    "supremacy_n24_f1": (REF / "circuits/supremacy_n24.qasm", 8, 1, ["--no-kat"]),
    "dnn_n25_f1": (REF / "circuits/dnn_n25.qasm", 8, 1, ["--no-kat"]),
    "supremacy_n26_f1": (REF / "circuits/supremacy_n26.qasm", 8, 1, ["--no-kat"]),
}
# sharded schedules of small circuits (trace + manifest only, committed): name -> (circuit, threads, fuse, world)
SHARDED = {
    "mix_n10_f1_w2": ("tests/circuits/mix_n10.qasm", 4, 1, 2),
    "mix_n10_f0_w2": ("tests/circuits/mix_n10.qasm", 4, 0, 2),
    "mix_n10_f3_w4": ("tests/circuits/mix_n10.qasm", 4, 3, 4),
    "qft_n8_f0_w2": ("tests/circuits/qft_n8.qasm", 4, 0, 2),
    "mix_n12_f1_w4": ("tests/circuits/mix_n12.qasm", 8, 1, 4),
    "mix_n12_f3_w8": ("tests/circuits/mix_n12.qasm", 8, 3, 8),
    "brick_n11_f3_w2": ("tests/circuits/brick_n11.qasm", 8, 3, 2),
    # the dependency-graph fusion (the schedule bench.py measures) under sharding
    "brick_n11_f4_w2": ("tests/circuits/brick_n11.qasm", 8, 4, 2),
    "mix_n12_f4_w4": ("tests/circuits/mix_n12.qasm", 8, 4, 4),
    "mix_n10_f4_w8": ("tests/circuits/mix_n10.qasm", 4, 4, 8),
    # compound operations are expanded in sharded mode (a SWAP inside one renames qubits mid-operation)
    "compound_n9_f0_w2": ("tests/circuits/compound_n9.qasm", 4, 0, 2),
    "compound_n9_f1_w4": ("tests/circuits/compound_n9.qasm", 4, 1, 4),
    "compound_n9_f4_w2": ("tests/circuits/compound_n9.qasm", 4, 4, 2),
    "compound_n9_f4_w8": ("tests/circuits/compound_n9.qasm", 4, 4, 8),
}
# traces only (host DD phase + fusion schedule, no reference array phase): name -> (circuit, fuse[, shards]);
# fuse 4 = dependency-graph fusion with the GPU cost model (the schedule bench.py measures)
TRACES = {
    "supremacy_n26_gpu": (REF / "circuits/supremacy_n26.qasm", 4),
    "supremacy_n26_gpu_w2": (REF / "circuits/supremacy_n26.qasm", 4, 2),
    "supremacy_n26_gpu_w4": (REF / "circuits/supremacy_n26.qasm", 4, 4),
    "supremacy_n26_gpu_w8": (REF / "circuits/supremacy_n26.qasm", 4, 8),
    "supremacy_n20_gpu_w2": (REF / "circuits/supremacy_n20.qasm", 4, 2),
    "knn_n25_f0_w2": (REF / "circuits/knn_n25.qasm", 0, 2),
    "synth_n26_w1": (ROOT / "bench_inputs/circuits/synth_n26.qasm", 4, 1),
    "synth_n26_w8": (ROOT / "bench_inputs/circuits/synth_n26.qasm", 4, 8),
    "synth_n30_w1": (ROOT / "bench_inputs/circuits/synth_n30.qasm", 4, 1),
    "synth_n30_w2": (ROOT / "bench_inputs/circuits/synth_n30.qasm", 4, 2),
    "synth_n30_w4": (ROOT / "bench_inputs/circuits/synth_n30.qasm", 4, 4),
    "synth_n30_w8": (ROOT / "bench_inputs/circuits/synth_n30.qasm", 4, 8),
    "synth_n34_w8": (ROOT / "bench_inputs/circuits/synth_n34.qasm", 4, 8),
    "synth_n20_w2": (ROOT / "bench_inputs/circuits/synth_n20.qasm", 4, 2),
    "knn_n31_f0_w1": (REF / "circuits/knn_n31.qasm", 0, 1),
    "knn_n31_f0_w2": (REF / "circuits/knn_n31.qasm", 0, 2),
    "knn_n31_f0_w4": (REF / "circuits/knn_n31.qasm", 0, 4),
    "knn_n31_f0_w8": (REF / "circuits/knn_n31.qasm", 0, 8),
    "supremacy_n26_ref": (REF / "circuits/supremacy_n26.qasm", 1),
    "supremacy_n24_gpu": (REF / "circuits/supremacy_n24.qasm", 4),
    "supremacy_n20_gpu": (REF / "circuits/supremacy_n20.qasm", 4),
    "dnn_n25_gpu": (REF / "circuits/dnn_n25.qasm", 4),
    "knn_n25_gpu": (REF / "circuits/knn_n25.qasm", 4),
    "swap_test_n25_gpu": (REF / "circuits/swap_test_n25.qasm", 4),
    "vqe_n16_gpu": (REF / "circuits/vqe_n16.qasm", 4),
}


def run_case(out_root: Path, name: str, circuit, threads: int, fuse: int, extra):
    out = out_root / name
    out.mkdir(parents=True, exist_ok=True)
    cmd = [str(DUMP), "--file", str(ROOT / circuit if not Path(circuit).is_absolute() else circuit), "--out", str(out),
           "-t", str(threads), "--fuse", str(fuse)] + list(extra)
    print(" ".join(cmd), flush=True)
    with open(out / "ref_dump.log", "w") as log:
        subprocess.run(cmd, check=True, stdout=log, stderr=subprocess.STDOUT)


def main(argv):
    what = argv[1] if len(argv) > 1 else "small"
    only = set(argv[2:])
    if not DUMP.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], check=True)
    if what == "small":
        for name, (c, t, f, extra) in SMALL.items():
            if not only or name in only:
                run_case(ROOT / "tests" / "golden", name, c, t, f, ["--full-state"] + extra)
    elif what in ("medium", "large"):
        table = MEDIUM if what == "medium" else LARGE
        for name, (c, t, f, extra) in table.items():
            if not only or name in only:
                run_case(ROOT / "oracle" / "_ref" / "golden", name, c, t, f, extra)
    elif what == "traces":
        for name, spec in TRACES.items():
            if not only or name in only:
                extra = ["--trace-fuse", str(spec[1]), "--no-ref"] + (["--world", str(spec[2])] if len(spec) > 2 else [])
                run_case(ROOT / "oracle" / "_ref" / "traces", name, spec[0], 8, 1, extra)
    elif what == "sharded":
        for name, (c, t, f, w) in SHARDED.items():
            if not only or name in only:
                run_case(ROOT / "tests" / "golden_sharded", name, c, t, 1, ["--trace-fuse", str(f), "--world", str(w), "--no-ref"])
    elif what == "samples":
        make_sample("supremacy_n26", after_switch=160)
    elif what == "circuits":
        # input data for the CLI on the GPU box (which has no reference tree); git-ignored
        import shutil
        dst = ROOT / "oracle" / "_ref" / "circuits"
        dst.mkdir(parents=True, exist_ok=True)
        for q in sorted((REF / "circuits").glob("*.qasm")):
            shutil.copyfile(q, dst / q.name)
    elif what == "publish":
        # the reference's sampled final amplitudes of the workloads bench.py runs: bench.py compares its own final
        # state with them at every GPU count (check.max_amp_err_vs_reference)
        import shutil
        dst = ROOT / "bench_inputs" / "samples"
        dst.mkdir(parents=True, exist_ok=True)
        for workload, case in (("supremacy_n26", "supremacy_n26_f1"), ("knn_n31_f0", "knn_n31_f0")):
            src = ROOT / "oracle" / "_ref" / "golden" / case
            if (src / "samples.bin").exists():
                shutil.copyfile(src / "samples.bin", dst / f"{workload}.samples.bin")
                shutil.copyfile(src / "manifest.json", dst / f"{workload}.manifest.json")
    else:
        raise SystemExit(__doc__)


def make_sample(name: str, after_switch: int):
    """First (switch op + 1 + after_switch) operations of a reference circuit, so the unmodified
    reference CLI switches at the same operation and then runs a bounded array phase."""
    import json
    man = json.loads((ROOT / "oracle" / "_ref" / "traces" / f"{name}_ref" / "manifest.json").read_text())
    keep = man["trace"]["switched_at_op"] + 1 + after_switch
    out_dir = ROOT / "oracle" / "_ref" / "circuits"
    out_dir.mkdir(parents=True, exist_ok=True)
    lines, ops = [], 0
    for line in (REF / "circuits" / f"{name}.qasm").read_text().splitlines():
        is_op = "q[" in line and not line.lstrip().startswith(("//", "qreg", "creg"))
        if is_op:
            if ops >= keep:
                continue
            ops += 1
        lines.append(line)
    (out_dir / f"{name}_sample.qasm").write_text("\n".join(lines) + "\n")
    (out_dir / f"{name}_sample.json").write_text(json.dumps({"source": f"{name}.qasm", "unitary_ops": ops,
                                                            "switched_at_op": man["trace"]["switched_at_op"]}))


if __name__ == "__main__":
    main(sys.argv)
