"""Runs build/flatdd_gpu (or, with `standalone` as second argument, build/flatdd_gpu_standalone) on every reference
circuit present under third_party/ref_install/circuits and prints one JSON line per circuit (the CLI's own statistics block).
usage: python tools/run_all_circuits.py [fuse] [standalone|dd] [time-gates]
(time-gates synchronises after every launch to report per-launch device time; array_phase_time is then not representative)"""
import json
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
standalone = len(sys.argv) > 2 and sys.argv[2] == "standalone"
CLI = ROOT / "build" / ("flatdd_gpu_standalone" if standalone else "flatdd_gpu")
fuse = sys.argv[1] if len(sys.argv) > 1 else ("2" if standalone else "4")
names = ["ghz_state_n23", "vqe_n16", "dnn_n16", "dnn_n20", "supremacy_n20", "supremacy_n24", "knn_n25", "swap_test_n25", "dnn_n25",
         "supremacy_n26", "adder_n28", "knn_n31"]
for name in names:
    circuit = ROOT / "third_party" / "ref_install" / "circuits" / f"{name}.qasm"
    if not circuit.exists():
        continue
    with tempfile.TemporaryDirectory() as tmp:
        cwd = Path(tmp) / "build" / "apps"
        cwd.mkdir(parents=True)
        (Path(tmp) / "log" / "results" / "time").mkdir(parents=True)
        (Path(tmp) / "log" / "results" / "state").mkdir(parents=True)
        t0 = time.perf_counter()
        # (round 1 ran knn_n31 per gate in the DD-driven binary: DD-level fusion of cswap chains was slow; the dense-block fusion is not)
        f = "0" if name.startswith("knn_n31") and not standalone and fuse in ("1", "2", "3", "5") else fuse
        res = subprocess.run([str(CLI), "--file", str(circuit), "-t", "16", "--fuse", f, "--quiet"] + (["--time-gates"] if "time-gates" in sys.argv[1:] else []), cwd=cwd, capture_output=True, text=True)
        wall = time.perf_counter() - t0
    if res.returncode != 0:
        print(json.dumps({"benchmark": name, "error": res.stderr[-300:]}))
        continue
    out = res.stdout
    stats = json.loads(out[out.rindex("\n{\n") + 1:])["statistics"]
    stats["wall_s"] = wall
    stats["fuse"] = int(f)
    stats["cli"] = CLI.name
    print(json.dumps(stats), flush=True)
