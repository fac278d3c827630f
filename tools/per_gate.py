"""Per-launch device times of a boundary trace with the structural facts of every gate.
usage: python tools/per_gate.py <trace-name> [out.csv] [key=value options...]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import find_trace  # noqa: E402
from flatdd_b200 import Context, load_library, read_trace  # noqa: E402


def main():
    name = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 and "=" not in sys.argv[2] else None
    opts = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
    lib = load_library()
    n, records = read_trace(find_trace(name))
    rows = []
    with Context(n) as ctx:
        for k, v in opts.items():
            ctx.set_option(k, int(v))
        gates = [(r, ctx.compile(r.dd)) for r in records if r.kind == 2]
        for rep in range(2):
            ctx.set_timing(False)
            ctx.convert(records[0].dd)
            ctx.set_timing(True)
            rows = []
            for i, (r, g) in enumerate(gates):
                ctx.apply_compiled(g)
                rows.append((i, r.n_original_gates, r.dd.n_nodes, g.info("max_paths"), g.info("max_sub_k"), g.info("upper_depth"),
                             g.info("upper_nodes"), g.info("sub_tables"), g.info("top_level"), g.info("tileable"), g.info("non_diag_upper"), g.info("uniform"),
                             g.info("block_targets"), g.info("block_context"), ctx.last_kernel_ms()))
    hdr = "idx,orig,nodes,paths,k,depth,upper,subs,top,tileable,ndu,uniform,block_k,block_ctx,ms"
    lines = [hdr] + [",".join(str(x) for x in r) for r in rows]
    if out:
        Path(out).parent.mkdir(parents=True, exist_ok=True)
        Path(out).write_text("\n".join(lines) + "\n")
    tot = sum(r[-1] for r in rows)
    print(f"{name}: {len(rows)} launches, total {tot:.2f} ms, mean {tot / len(rows):.3f} ms, GB/s {32 * (1 << n) / (tot / len(rows) * 1e-3) / 1e9:.0f}  opts={opts}")
    # aggregate by (paths, k)
    agg = {}
    for r in rows:
        agg.setdefault((r[3], r[4]), []).append(r[-1])
    for key in sorted(agg):
        v = agg[key]
        print(f"  paths={key[0]:3d} k={key[1]:2d}: {len(v):4d} launches, mean {sum(v) / len(v):.3f} ms, min {min(v):.3f}, max {max(v):.3f}")
    uni = [r[-1] for r in rows if r[11] == 1]
    blk = [r[-1] for r in rows if r[12] >= 0]
    print(f"  dense blocks: {len(blk)} of {len(rows)}" + (f", mean {sum(blk) / len(blk):.3f} ms" if blk else ""))
    print(f"  uniform gates: {len(uni)} of {len(rows)}" + (f", mean {sum(uni) / len(uni):.3f} ms" if uni else ""))


if __name__ == "__main__":
    main()
