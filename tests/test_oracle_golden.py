"""CPU tests: the C restatement (oracle/flat_oracle.c) against golden vectors written by the
compiled reference itself (oracle/ref_dump.cpp -> tests/golden/).  Bit-exact unless stated."""
import numpy as np
import pytest

from flatdd_b200 import read_flat, read_trace
from oracle import pyoracle
from tests import golden_util as G

CASES = G.cases()


def test_golden_cases_exist():
    assert len(CASES) >= 12


@pytest.mark.parametrize("case", CASES)
def test_convert_walk_bit_exact(case):
    """oracle_convert == reference getValueByPathPar on every amplitude (SwitchPackage.hpp:3605-3634)."""
    dd = read_flat(G.GOLDEN / case / "kat_convert_dd.bin")
    re, im = pyoracle.convert(dd)
    assert np.array_equal(re, G.f64(case, "kat_convert_walk_re.f64"))
    assert np.array_equal(im, G.f64(case, "kat_convert_walk_im.f64"))


@pytest.mark.parametrize("case", CASES)
def test_convert_switch1_bit_exact(case):
    """oracle_convert_switch1 == reference getVectorFromDDSwitch1 incl. the regularity shortcut."""
    m = G.manifest(case)
    dd = read_flat(G.GOLDEN / case / "kat_convert_dd.bin")
    n_thread_exp = int(np.log2(m["threads"]))
    re, im = pyoracle.convert_switch1(dd, n_thread_exp)
    assert np.array_equal(re, G.f64(case, "kat_convert_switch1_re.f64"))
    assert np.array_equal(im, G.f64(case, "kat_convert_switch1_im.f64"))
    # and the shortcut stays within rounding of the exact walk
    wr, wi = pyoracle.convert(dd)
    assert G.max_amp_err(re, im, wr, wi) < 1e-15


@pytest.mark.parametrize("case", CASES)
def test_dd_size(case):
    m = G.manifest(case)
    kat = [k for k in m["kats"] if k["kind"] == "convert"][0]
    dd = read_flat(G.GOLDEN / case / kat["dd"])
    assert pyoracle.dd_size(dd) == kat["dd_size"]


def _dmavm_kats():
    out = []
    for case in CASES:
        for k in G.manifest(case)["kats"]:
            if k["kind"] == "dmavm":
                out.append((case, k["stem"]))
    return out


@pytest.mark.parametrize("case,stem", _dmavm_kats())
def test_dmavm_bit_exact(case, stem):
    """oracle_dmavm == reference DDArrMultiplyIP on a seeded random state (SwitchPackage.hpp:1897-2261)."""
    gate = read_flat(G.GOLDEN / case / f"{stem}_dd.bin")
    zr, zi = pyoracle.dmavm(gate, G.f64(case, f"{stem}_y_re.f64"), G.f64(case, f"{stem}_y_im.f64"))
    assert np.array_equal(zr, G.f64(case, f"{stem}_z_re.f64"))
    assert np.array_equal(zi, G.f64(case, f"{stem}_z_im.f64"))


@pytest.mark.parametrize("case,stem", _dmavm_kats())
def test_cost_model(case, stem):
    """nnz, DMAVMACStatIP, DMAVMACStatOP1 and size() match the reference (SwitchPackage.hpp:3006-3311)."""
    m = G.manifest(case)
    k = [x for x in m["kats"] if x.get("stem") == stem][0]
    gate = read_flat(G.GOLDEN / case / f"{stem}_dd.bin")
    t = int(np.log2(k["threads"]))
    assert pyoracle.mac_count(gate) == k["nnz"]
    assert pyoracle.cost_ip(gate, t) == k["cost_ip"]
    assert pyoracle.cost_op1(gate, t) == k["cost_op1"]
    assert pyoracle.dd_size(gate) == k["dd_size"]


@pytest.mark.parametrize("case", CASES)
def test_trace_replay_matches_reference_final_state(case):
    """The product's host driver (GpuSwitchSimulator) recorded a boundary trace; replaying it on
    the oracle must reproduce the reference's own SwitchSimulator::simulate() final state.
    fuse 0 and --no_cache runs use only DDArrMultiplyIP and the exact walk, so they are
    bit-identical unless the regularity shortcut fired; cache (OP) runs agree to rounding."""
    m = G.manifest(case)
    n, records = read_trace(G.GOLDEN / case / "trace.bin")
    assert n == m["n_qubits"] and len(records) == m["trace"]["records"]
    assert m["trace"]["switched"] == m["reference"]["switched"]
    if m["reference"]["switched"]:
        assert m["trace"]["launches"] <= m["reference"]["array_phase_launches"]  # identity gates are not launched
    re, im = pyoracle.replay_trace(records)
    fr, fi = G.final_state(case)
    err = G.max_amp_err(re, im, fr, fi)
    assert err < 1e-14, err
    assert 1.0 - G.fidelity(re, im, fr, fi) < 1e-13


def test_switch_rule():
    """EMA rule (src/SwitchSimulator.cpp:97,163-165,181): switch iff old EMA * thr < size."""
    # EMA_0 = 4; sizes 5 -> ema 4.1; 9 > 8.2 -> switch at index 1
    assert pyoracle.switch_index(4, [5, 9, 100]) == 1
    assert pyoracle.switch_index(4, [5, 8, 8, 8]) == -1
    assert pyoracle.switch_index(10, [21]) == 0
    assert pyoracle.switch_index(10, [20, 20, 20]) == -1
