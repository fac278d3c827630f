"""CPU tests of the product library's host-only entry points (no GPU needed): it loads, exports
every symbol include/flatdd_b200.h declares, reproduces the reference's cost model on the golden
gates, rejects malformed tables, and fails loudly without a CUDA device."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from flatdd_b200 import FlatDDError, load_library, read_flat
from flatdd_b200.capi import EXPORTS
from tests import golden_util as G

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "flatdd_b200.h").read_text()
    declared = set(re.findall(r"\b(fdd_[a-z0-9_]+)\s*\(", header))
    declared -= {"fdd_ctx", "fdd_gate", "fdd_vecdd", "fdd_matdd"}
    lib = load_library().lib
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)


def _dmavm_kats():
    out = []
    for case in G.cases():
        for k in G.manifest(case)["kats"]:
            if k["kind"] == "dmavm":
                out.append((case, k))
    return out


@pytest.mark.parametrize("case,kat", _dmavm_kats(), ids=lambda x: x if isinstance(x, str) else x["stem"])
def test_cost_model_matches_reference(case, kat):
    """fdd_mac_count / fdd_cost_ip / fdd_cost_op1 == DMAVMACCountIP / DMAVMACStatIP / DMAVMACStatOP1 of the reference."""
    lib = load_library()
    gate = read_flat(G.GOLDEN / case / f"{kat['stem']}_dd.bin")
    t = int(np.log2(kat["threads"]))
    assert lib.mac_count(gate) == kat["nnz"]
    assert lib.cost_ip(gate, t) == kat["cost_ip"]
    assert lib.cost_op1(gate, t) == kat["cost_op1"]
    assert lib.cost_gpu(gate) > 0
    assert lib.matdd_info(gate, "nnz") == kat["nnz"]


def test_malformed_tables_are_rejected():
    lib = load_library()
    gate = read_flat(G.GOLDEN / "mix_n10_f0" / "kat_gate0_dd.bin")
    bad = read_flat(G.GOLDEN / "mix_n10_f0" / "kat_gate0_dd.bin")
    bad.child[bad.root, 0] = gate.n_nodes + 5
    with pytest.raises(FlatDDError):
        lib.mac_count(bad)
    bad = read_flat(G.GOLDEN / "mix_n10_f0" / "kat_gate0_dd.bin")
    bad.level[bad.root] = 3
    with pytest.raises(FlatDDError):
        lib.cost_ip(bad, 2)
    bad = read_flat(G.GOLDEN / "mix_n10_f0" / "kat_gate0_dd.bin")
    bad.weight[bad.root, 0, 0] = np.nan
    with pytest.raises(FlatDDError):
        lib.cost_gpu(bad)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails with FDD_ERR_CUDA; with one this test is moot."""
    lib = load_library()
    try:
        n = lib.device_count()
    except FlatDDError as exc:
        assert exc.code == -2
        from flatdd_b200 import Context
        with pytest.raises(FlatDDError):
            Context(4)
        return
    assert n >= 1


def test_comm_unique_id_has_nccl_size():
    assert len(load_library().comm_unique_id()) == 128


def test_tile_facts_of_compiled_gates():
    """Host analysis that selects the kernel path: tile bits, uniformity and the context bits of the tensor-core path."""
    from tests import dd_builder as B
    lib = load_library()
    rng = np.random.default_rng(3)
    dense = B.gate_dd(14, [13, 9, 6, 5], B.random_unitary(4, rng))
    assert lib.matdd_info(dense, "tileable") == 1 and lib.matdd_info(dense, "uniform") == 1
    assert lib.matdd_info(dense, "sub_tile_bits") == 4 and lib.matdd_info(dense, "max_paths") == 16
    assert lib.matdd_info(dense, "context_bits") == 0 and lib.matdd_info(dense, "max_sub_k") == 1
    # controls on qubits 12 and 7 (upper, outside the tile), dense on 14, 6, 9, 11: two context bits
    ctrl = B.gate_dd(15, [12, 7, 14, 6, 9, 11], B.controlled(B.random_unitary(4, rng), 2))
    assert lib.matdd_info(ctrl, "uniform") == 0 and lib.matdd_info(ctrl, "context_bits") == 2
    assert lib.matdd_info(ctrl, "sub_tile_bits") == 4 and lib.matdd_info(ctrl, "non_diag_upper") == 4
    # a control on a lane qubit is part of the sub table, not a context bit
    lane = B.gate_dd(15, [2, 14, 6, 9, 11], B.controlled(B.random_unitary(4, rng), 1))
    assert lib.matdd_info(lane, "context_bits") == 0 and lib.matdd_info(lane, "sub_tables") > 1
    # the tensor-core classes are priced near one pass, a sub table per path clearly above
    mem_ns = 32.0 * 2 ** 14 / 6500.0
    assert (lib.cost_gpu(dense) - 3000.0) / mem_ns < 1.4
    assert lib.cost_gpu(lane) > lib.cost_gpu(dense)
