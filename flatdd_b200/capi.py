"""ctypes binding of include/flatdd_b200.h.  Every failure of the library raises FlatDDError;
a missing library is an ImportError-like hard failure (there is no Python/CPU fallback)."""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

from .flat import FlatDD, _CVecDD

_PKG = Path(__file__).resolve().parent


def library_path() -> Path:
    override = os.environ.get("FLATDD_B200_LIB")  # experiments only: an alternative build of the same library
    return Path(override) if override else _PKG / "libflatdd_b200.so"


class FlatDDError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"flatdd_b200 error {code}: {message}")
        self.code = code


EXPORTS = [
    "fdd_version", "fdd_last_error", "fdd_device_count", "fdd_create", "fdd_create_sharded", "fdd_destroy",
    "fdd_n_qubits", "fdd_n_local_qubits", "fdd_synchronize", "fdd_set_option", "fdd_get_option", "fdd_comm_unique_id", "fdd_comm_init",
    "fdd_exchange_qubits", "fdd_apply_many_exchange", "fdd_gate_apply_many_exchange", "fdd_relabel_qubits", "fdd_barrier",
    "fdd_convert", "fdd_apply", "fdd_apply_many", "fdd_block_from_matdd", "fdd_gate_compile", "fdd_gate_apply", "fdd_gate_apply_many", "fdd_gate_free", "fdd_gate_info",
    "fdd_ddarr_multiply", "fdd_mac_count", "fdd_cost_ip", "fdd_cost_op1", "fdd_cost_gpu", "fdd_matdd_info", "fdd_get_state",
    "fdd_set_state", "fdd_set_zero_state", "fdd_get_amplitudes", "fdd_get_amplitudes_at", "fdd_norm2", "fdd_sample", "fdd_state_device_ptr",
    "fdd_get_permutation", "fdd_canonicalize", "fdd_last_kernel_ms", "fdd_set_timing", "fdd_launch_count", "fdd_stream",
]


class Library:
    def __init__(self, path=None):
        path = Path(path) if path is not None else library_path()
        if not path.exists():
            raise FileNotFoundError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(flatdd_b200 has no CPU fallback)")
        self.path = path
        self.lib = ctypes.CDLL(str(path))
        L = self.lib
        vp, dp, i32 = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_int
        ddp = ctypes.POINTER(_CVecDD)
        L.fdd_version.restype = ctypes.c_char_p
        L.fdd_last_error.restype = ctypes.c_char_p
        L.fdd_device_count.argtypes = [ctypes.POINTER(i32)]
        L.fdd_create.argtypes = [i32, i32, ctypes.POINTER(vp)]
        L.fdd_create_sharded.argtypes = [i32, i32, i32, i32, ctypes.POINTER(vp)]
        L.fdd_destroy.argtypes = [vp]
        L.fdd_n_qubits.argtypes = [vp]
        L.fdd_n_local_qubits.argtypes = [vp]
        L.fdd_synchronize.argtypes = [vp]
        L.fdd_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_long]
        L.fdd_get_option.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_long)]
        L.fdd_comm_unique_id.argtypes = [vp]
        L.fdd_comm_init.argtypes = [vp, vp]
        L.fdd_exchange_qubits.argtypes = [vp, i32, i32, i32]
        L.fdd_relabel_qubits.argtypes = [vp, i32, i32]
        L.fdd_barrier.argtypes = [vp]
        L.fdd_convert.argtypes = [vp, ddp]
        L.fdd_apply.argtypes = [vp, ddp]
        L.fdd_apply_many.argtypes = [vp, ddp, i32]
        L.fdd_gate_compile.argtypes = [vp, ddp, ctypes.POINTER(vp)]
        L.fdd_gate_apply.argtypes = [vp, vp]
        L.fdd_gate_apply_many.argtypes = [vp, ctypes.POINTER(vp), i32]
        L.fdd_gate_free.argtypes = [vp]
        L.fdd_gate_info.argtypes = [vp, ctypes.c_char_p]
        L.fdd_gate_info.restype = ctypes.c_long
        L.fdd_ddarr_multiply.argtypes = [ddp, dp, dp, dp, dp, ctypes.c_size_t, i32]
        L.fdd_mac_count.argtypes = [ddp, ctypes.POINTER(ctypes.c_uint64)]
        L.fdd_cost_ip.argtypes = [ddp, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint64)]
        L.fdd_cost_op1.argtypes = [ddp, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint64)]
        L.fdd_cost_gpu.argtypes = [ddp, ctypes.c_double, ctypes.c_double, dp]
        L.fdd_matdd_info.argtypes = [ddp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_long)]
        L.fdd_get_state.argtypes = [vp, dp, dp]
        L.fdd_set_state.argtypes = [vp, dp, dp]
        L.fdd_set_zero_state.argtypes = [vp]
        L.fdd_get_amplitudes.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64, dp]
        L.fdd_get_amplitudes_at.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64, dp]
        L.fdd_norm2.argtypes = [vp, dp]
        L.fdd_sample.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
        L.fdd_state_device_ptr.argtypes = [vp, ctypes.POINTER(vp)]
        L.fdd_get_permutation.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
        L.fdd_canonicalize.argtypes = [vp]
        L.fdd_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
        L.fdd_set_timing.argtypes = [vp, i32]
        L.fdd_launch_count.argtypes = [vp]
        L.fdd_launch_count.restype = ctypes.c_uint64
        L.fdd_stream.argtypes = [vp, ctypes.POINTER(vp)]

    def check(self, rc: int) -> None:
        if rc != 0:
            raise FlatDDError(rc, self.lib.fdd_last_error().decode(errors="replace"))

    def version(self) -> str:
        return self.lib.fdd_version().decode()

    def device_count(self) -> int:
        n = ctypes.c_int(0)
        self.check(self.lib.fdd_device_count(ctypes.byref(n)))
        return n.value

    def comm_unique_id(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        self.check(self.lib.fdd_comm_unique_id(buf))
        return buf.raw

    # ---- host-only cost model ---------------------------------------------------------------
    def mac_count(self, gate: FlatDD) -> int:
        out = ctypes.c_uint64(0)
        c = gate.as_c()
        self.check(self.lib.fdd_mac_count(ctypes.byref(c), ctypes.byref(out)))
        return out.value

    def cost_ip(self, gate: FlatDD, n_thread_exp: int) -> int:
        out = ctypes.c_uint64(0)
        c = gate.as_c()
        self.check(self.lib.fdd_cost_ip(ctypes.byref(c), n_thread_exp, ctypes.byref(out)))
        return out.value

    def cost_op1(self, gate: FlatDD, n_thread_exp: int) -> int:
        out = ctypes.c_uint64(0)
        c = gate.as_c()
        self.check(self.lib.fdd_cost_op1(ctypes.byref(c), n_thread_exp, ctypes.byref(out)))
        return out.value

    def cost_gpu(self, gate: FlatDD, hbm_gbs: float = 6500.0, fp64_gflops: float = 30000.0) -> float:
        out = ctypes.c_double(0)
        c = gate.as_c()
        self.check(self.lib.fdd_cost_gpu(ctypes.byref(c), hbm_gbs, fp64_gflops, ctypes.byref(out)))
        return out.value

    def matdd_info(self, gate: FlatDD, key: str) -> int:
        out = ctypes.c_long(0)
        c = gate.as_c()
        self.check(self.lib.fdd_matdd_info(ctypes.byref(c), key.encode(), ctypes.byref(out)))
        return out.value

    def ddarr_multiply(self, gate: FlatDD, y_re: np.ndarray, y_im: np.ndarray, device: int = 0):
        """Literal DDArrMultiplyIP drop-in on host SoA arrays."""
        y_re = np.ascontiguousarray(y_re, dtype=np.float64)
        y_im = np.ascontiguousarray(y_im, dtype=np.float64)
        z_re = np.empty_like(y_re)
        z_im = np.empty_like(y_im)
        c = gate.as_c()
        dp = ctypes.POINTER(ctypes.c_double)
        self.check(self.lib.fdd_ddarr_multiply(ctypes.byref(c), y_re.ctypes.data_as(dp), y_im.ctypes.data_as(dp),
                                               z_re.ctypes.data_as(dp), z_im.ctypes.data_as(dp), y_re.size, device))
        return z_re, z_im


_LIB = None


def load_library(path=None) -> Library:
    global _LIB
    if _LIB is None or path is not None:
        _LIB = Library(path)
    return _LIB


class CompiledGate:
    def __init__(self, ctx: "Context", handle):
        self._ctx = ctx
        self._h = handle

    def info(self, key: str) -> int:
        return int(self._ctx.L.lib.fdd_gate_info(self._h, key.encode()))

    def free(self):
        if self._h:
            self._ctx.L.lib.fdd_gate_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One simulator state on one GPU (or one shard of a distributed state)."""

    def __init__(self, n_qubits: int, device: int = 0, rank: int = 0, world_size: int = 1, library: Library | None = None):
        self.L = library or load_library()
        h = ctypes.c_void_p()
        self.L.check(self.L.lib.fdd_create_sharded(n_qubits, device, rank, world_size, ctypes.byref(h)))
        self._h = h
        self.n_qubits = n_qubits
        self.n_local = int(self.L.lib.fdd_n_local_qubits(h))

    def close(self):
        if self._h:
            self.L.lib.fdd_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU -------------------------------------------------------------------------------
    def comm_init(self, unique_id: bytes):
        assert len(unique_id) == 128
        buf = ctypes.create_string_buffer(unique_id, 128)
        self.L.check(self.L.lib.fdd_comm_init(self._h, buf))

    def exchange_qubits(self, global_physical_bit: int, local_physical_bit: int, method: int = 0):
        self.L.check(self.L.lib.fdd_exchange_qubits(self._h, global_physical_bit, local_physical_bit, method))

    def relabel_qubits(self, physical_bit_a: int, physical_bit_b: int):
        self.L.check(self.L.lib.fdd_relabel_qubits(self._h, physical_bit_a, physical_bit_b))

    def barrier(self):
        self.L.check(self.L.lib.fdd_barrier(self._h))

    def permutation(self) -> np.ndarray:
        out = np.zeros(self.n_qubits, dtype=np.int32)
        self.L.check(self.L.lib.fdd_get_permutation(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return out

    def canonicalize(self):
        self.L.check(self.L.lib.fdd_canonicalize(self._h))

    def set_option(self, key: str, value: int):
        self.L.check(self.L.lib.fdd_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = ctypes.c_long(0)
        self.L.check(self.L.lib.fdd_get_option(self._h, key.encode(), ctypes.byref(v)))
        return int(v.value)

    def synchronize(self):
        self.L.check(self.L.lib.fdd_synchronize(self._h))

    def convert(self, dd: FlatDD):
        c = dd.as_c()
        self.L.check(self.L.lib.fdd_convert(self._h, ctypes.byref(c)))

    def apply(self, gate: FlatDD):
        c = gate.as_c()
        self.L.check(self.L.lib.fdd_apply(self._h, ctypes.byref(c)))

    def apply_many(self, gates):
        """Host tables of several gates in one boundary call: dense blocks among them share passes over the state."""
        gates = list(gates)
        if not gates:
            return
        arr = (type(gates[0].as_c()) * len(gates))(*[g.as_c() for g in gates])
        self.L.check(self.L.lib.fdd_apply_many(self._h, arr, len(gates)))

    def apply_many_exchange(self, gates, global_physical_bit: int, local_physical_bit: int):
        """A stretch of host tables and the exchange that follows it in one boundary call (the library fuses them when it can)."""
        gates = list(gates)
        arr = (type(gates[0].as_c()) * len(gates))(*[g.as_c() for g in gates]) if gates else None
        self.L.check(self.L.lib.fdd_apply_many_exchange(self._h, arr, len(gates), int(global_physical_bit), int(local_physical_bit)))

    def apply_compiled_many_exchange(self, gates, global_physical_bit: int, local_physical_bit: int):
        arr = (ctypes.c_void_p * len(gates))(*[g._h for g in gates]) if gates else None
        self.L.check(self.L.lib.fdd_gate_apply_many_exchange(self._h, arr, len(gates), int(global_physical_bit), int(local_physical_bit)))

    def compile(self, gate: FlatDD) -> CompiledGate:
        c = gate.as_c()
        h = ctypes.c_void_p()
        self.L.check(self.L.lib.fdd_gate_compile(self._h, ctypes.byref(c), ctypes.byref(h)))
        return CompiledGate(self, h)

    def apply_compiled(self, gate: CompiledGate):
        self.L.check(self.L.lib.fdd_gate_apply(self._h, gate._h))

    def apply_compiled_many(self, gates):
        """One C call for a run of compiled gates (a schedule segment between two exchanges)."""
        arr = (ctypes.c_void_p * len(gates))(*[g._h for g in gates])
        self.L.check(self.L.lib.fdd_gate_apply_many(self._h, arr, len(gates)))

    def set_zero_state(self):
        self.L.check(self.L.lib.fdd_set_zero_state(self._h))

    def set_state(self, re: np.ndarray, im: np.ndarray):
        re = np.ascontiguousarray(re, dtype=np.float64)
        im = np.ascontiguousarray(im, dtype=np.float64)
        assert re.size == im.size == 1 << self.n_local
        dp = ctypes.POINTER(ctypes.c_double)
        self.L.check(self.L.lib.fdd_set_state(self._h, re.ctypes.data_as(dp), im.ctypes.data_as(dp)))

    def get_state(self, out_re: np.ndarray | None = None, out_im: np.ndarray | None = None):
        dim = 1 << self.n_local
        re = out_re if out_re is not None else np.empty(dim, dtype=np.float64)
        im = out_im if out_im is not None else np.empty(dim, dtype=np.float64)
        dp = ctypes.POINTER(ctypes.c_double)
        self.L.check(self.L.lib.fdd_get_state(self._h, re.ctypes.data_as(dp), im.ctypes.data_as(dp)))
        return re, im

    def get_state_raw(self, re_ptr: int, im_ptr: int):
        """Same, into caller-owned (e.g. pinned) memory given as addresses."""
        dp = ctypes.POINTER(ctypes.c_double)
        self.L.check(self.L.lib.fdd_get_state(self._h, ctypes.cast(re_ptr, dp), ctypes.cast(im_ptr, dp)))

    def get_amplitudes(self, first: int, count: int) -> np.ndarray:
        out = np.empty(2 * count, dtype=np.float64)
        self.L.check(self.L.lib.fdd_get_amplitudes(self._h, first, count, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return out.view(np.complex128)

    def get_amplitudes_at(self, local_indices) -> np.ndarray:
        """Amplitudes at arbitrary local indices, gathered on the device."""
        idx = np.ascontiguousarray(local_indices, dtype=np.uint64)
        out = np.empty(2 * idx.size, dtype=np.float64)
        self.L.check(self.L.lib.fdd_get_amplitudes_at(self._h, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), idx.size,
                                                      out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return out.view(np.complex128)

    def norm2(self) -> float:
        out = ctypes.c_double(0)
        self.L.check(self.L.lib.fdd_norm2(self._h, ctypes.byref(out)))
        return out.value

    def sample(self, n_shots: int, seed: int = 0) -> np.ndarray:
        """Local physical indices of n_shots basis states drawn with probability |amp|^2 / shard norm."""
        out = np.zeros(n_shots, dtype=np.uint64)
        self.L.check(self.L.lib.fdd_sample(self._h, n_shots, seed, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
        return out

    def set_timing(self, enabled: bool):
        self.L.check(self.L.lib.fdd_set_timing(self._h, 1 if enabled else 0))

    def last_kernel_ms(self) -> float:
        out = ctypes.c_float(0)
        self.L.check(self.L.lib.fdd_last_kernel_ms(self._h, ctypes.byref(out)))
        return out.value

    def launch_count(self) -> int:
        return int(self.L.lib.fdd_launch_count(self._h))

    def stream(self) -> int:
        out = ctypes.c_void_p()
        self.L.check(self.L.lib.fdd_stream(self._h, ctypes.byref(out)))
        return out.value or 0

    def state_device_ptr(self) -> int:
        out = ctypes.c_void_p()
        self.L.check(self.L.lib.fdd_state_device_ptr(self._h, ctypes.byref(out)))
        return out.value or 0
