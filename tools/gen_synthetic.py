"""Synthetic supremacy-style circuit (input 5), emitted as OpenQASM 2 so the
same parser reads it.  Per layer: a random gate from {rx(pi/2), ry(pi/2), u3(theta,phi,lambda)} on
every qubit, a brick pattern of cx over a 2-D-grid-like pairing (alternating horizontal / vertical
neighbours on a width-w grid), and two (three for n > 30) extra random one-qubit gates on each of the three top
qubits so that more than 20 % of the gates touch the qubits that are global on 8 GPUs.
usage: python tools/gen_synthetic.py <n_qubits> <depth> <out.qasm> [seed]"""
import math
import random
import sys


def generate(n: int, depth: int, seed: int = 20241017):
    rng = random.Random(seed * 1000003 + n)
    width = max(2, int(round(math.sqrt(n))))
    lines = ["// synthetic supremacy-style circuit: tools/gen_synthetic.py", "OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{n}];"]

    def one_qubit(q):
        kind = rng.randrange(3)
        if kind == 0:
            lines.append(f"rx(pi*0.5) q[{q}];")
        elif kind == 1:
            lines.append(f"ry(pi*0.5) q[{q}];")
        else:
            lines.append(f"u3({rng.uniform(0, 2 * math.pi):.12f},{rng.uniform(0, 2 * math.pi):.12f},{rng.uniform(0, 2 * math.pi):.12f}) q[{q}];")

    total = touching_top = 0
    top = set(range(n - 3, n))
    for layer in range(depth):
        for q in range(n):
            one_qubit(q)
            total += 1
            touching_top += q in top
        pairs = []
        if layer % 2 == 0:  # horizontal neighbours
            for q in range(n):
                if q % width != width - 1 and q + 1 < n and (q % width) % 2 == (layer // 2) % 2:
                    pairs.append((q, q + 1))
        else:  # vertical neighbours
            for q in range(n - width):
                if (q // width) % 2 == (layer // 2) % 2:
                    pairs.append((q, q + width))
        for a, b in pairs:
            lines.append(f"cx q[{a}],q[{b}];")
            total += 1
            touching_top += (a in top) or (b in top)
        for q in sorted(top):
            for _ in range(2 if n <= 30 else 3):
                one_qubit(q)
                total += 1
                touching_top += 1
    return lines, total, touching_top


if __name__ == "__main__":
    n, depth, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 20241017
    lines, total, touching = generate(n, depth, seed)
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"{out}: {n} qubits, depth {depth}, {total} gates, {100.0 * touching / total:.1f}% touch the top 3 qubits")
