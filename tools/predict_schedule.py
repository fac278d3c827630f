"""Predicts the array-phase time of a boundary trace from measured per-class launch times (host only, no GPU).
usage: python tools/predict_schedule.py <trace.bin | trace-name> <per_gate.csv> [<per_gate.csv> ...]
The class of a launch is (max_sub_k, max_paths, sub_tables, non_diag_upper, uniform), as in tools/per_gate.py."""
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import find_trace  # noqa: E402
from flatdd_b200 import load_library, read_trace  # noqa: E402


def main():
    src = Path(sys.argv[1])
    n, records = read_trace(src if src.exists() else find_trace(sys.argv[1]))
    table = defaultdict(list)
    for csv in sys.argv[2:]:
        lines = Path(csv).read_text().splitlines()
        hdr = lines[0].split(",")
        for ln in lines[1:]:
            r = dict(zip(hdr, ln.split(",")))
            table[(int(r["k"]), int(r["paths"]), int(r["subs"]), int(r["ndu"]), int(r["uniform"]))].append(float(r["ms"]))
    mean = {k: sum(v) / len(v) for k, v in table.items()}
    lib = load_library()
    total, unknown, ops = 0.0, defaultdict(int), 0
    classes = defaultdict(int)
    gates = [r for r in records if r.kind == 2]
    for r in gates:
        key = tuple(lib.matdd_info(r.dd, k) for k in ("max_sub_k", "max_paths", "sub_tables", "non_diag_upper", "uniform"))
        classes[key] += 1
        ops += r.n_original_gates
        if key in mean:
            total += mean[key]
        else:
            unknown[key] += 1
            # nearest known class with at least as many paths and as wide a sub table, else 0.75 ms
            cand = [v for k, v in mean.items() if k[0] >= key[0] and k[1] >= key[1]]
            total += min(cand) if cand else 0.75
    print(f"{src}: n={n}, {len(gates)} launches for {ops} ops, predicted {total:.2f} ms, mean {total / max(1, len(gates)):.3f} ms"
          f" ({32 * (1 << n) / (total / max(1, len(gates)) * 1e-3) / 1e9:.0f} GB/s per launch)")
    for key, cnt in sorted(classes.items(), key=lambda kv: -kv[1]):
        print(f"   {key}: {cnt:3d} launches" + (f", {mean[key]:.3f} ms each" if key in mean else "  (class not measured)"))


if __name__ == "__main__":
    main()
