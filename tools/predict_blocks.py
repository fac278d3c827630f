"""Predicts the array-phase time of a dense-block schedule from measured pass times (no GPU): the library's greedy packing of
consecutive blocks into passes is replayed on the trace and every pass is priced as
    max(0.36, 0.10 + 0.33 * (16x16 blocks) + 0.20 * (8x8 blocks))  ms at n = 26   (B200, round-2 measurements, profiles/r02_*)
usage: python tools/predict_blocks.py <trace.bin> [max blocks per pass] [max upper targets per pass]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flatdd_b200 import load_library, read_trace  # noqa: E402


def predict(path, max_blocks=4, max_upper=7, verbose=False):
    lib = load_library()
    n, records = read_trace(path)
    scale = 2.0 ** (n - 26)
    gates = [r for r in records if r.kind == 2]
    passes, cur_upper, cur = [], set(), []
    other = 0
    for r in gates:
        mask = lib.matdd_info(r.dd, "non_diag_mask")
        targets = [q for q in range(n) if (mask >> q) & 1]
        if len(targets) > 4:
            if cur:
                passes.append(cur)
            cur, cur_upper = [], set()
            other += 1
            continue
        k = 4 if len(targets) == 4 else 3
        upper = {q for q in targets if q >= 5}
        if cur and (len(cur) >= max_blocks or len(cur_upper | upper) > max_upper):
            passes.append(cur)
            cur, cur_upper = [], set()
        cur.append(k)
        cur_upper |= upper
    if cur:
        passes.append(cur)
    total = 0.0
    for p in passes:
        t = max(0.36, 0.10 + 0.33 * p.count(4) + 0.20 * p.count(3)) if len(p) > 1 else (0.385 if p[0] == 4 else 0.36)
        total += t * scale
    total += other * 0.7 * scale
    if verbose:
        print("passes:", " ".join("".join(map(str, p)) for p in passes))
    return {"gates": len(gates), "passes": len(passes), "other": other, "k4": sum(p.count(4) for p in passes), "k3": sum(p.count(3) for p in passes), "ms": round(total, 2)}


if __name__ == "__main__":
    mb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    mu = int(sys.argv[3]) if len(sys.argv) > 3 else 7
    print(predict(sys.argv[1], mb, mu, verbose=True))
