"""ctypes binding of libqcsim_b200.so (the C ABI declared in include/qcsim_b200.h).

There is no CPU fallback: if the CUDA extension is missing or no device is present, the product
path raises.  Nothing here imports anything from oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqcsim_b200.so")

OK = 0
ERR_QUBIT_TOO_HIGH = -1
ERR_CTRL_TOO_HIGH = -2
ERR_SAME_QUBITS = -3
ERR_BAD_ARG = -4
ERR_BAD_STATE = -5
ERR_CUDA = -6
ERR_NCCL = -7
ERR_OOM = -8
ERR_UNSUPPORTED = -9


class QcsimError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"qcsim_b200 error {code}: {msg}")
        self.code = code


class GateStruct(C.Structure):
    """struct qcsim_gate"""

    _fields_ = [("nq", C.c_int32), ("flags", C.c_int32), ("q", C.c_uint64), ("c1", C.c_uint64), ("c2", C.c_uint64),
                ("m", C.c_double * 128)]


class CircuitGateStruct(C.Structure):
    """struct qcsim_circuit_gate (one record of a circuit file)"""

    _fields_ = [("nq", C.c_int32), ("flags", C.c_int32), ("gate_id", C.c_int32), ("reserved", C.c_int32), ("q", C.c_uint64),
                ("c1", C.c_uint64), ("c2", C.c_uint64), ("params", C.c_double * 4), ("m", C.c_double * 128)]


class Stats(C.Structure):
    """struct qcsim_stats"""

    _fields_ = [("gates_applied", C.c_uint64), ("kernel_launches", C.c_uint64), ("state_passes", C.c_uint64),
                ("bytes_moved", C.c_uint64), ("exchange_calls", C.c_uint64), ("exchange_bytes", C.c_uint64),
                ("exchange_ms", C.c_double), ("fused_rounds", C.c_uint64), ("fused_ops", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_P = C.c_void_p
_U64 = C.c_uint64
_DP = C.POINTER(C.c_double)
_U64P = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); every symbol include/qcsim_b200.h declares
SIGNATURES = {
    "qcsim_last_error": (C.c_char_p, []),
    "qcsim_abi_version": (C.c_int, []),
    "qcsim_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "qcsim_sv_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    "qcsim_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "qcsim_sv_create_sharded": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "qcsim_sv_create_multi": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "qcsim_sv_destroy": (C.c_int, [_P]),
    "qcsim_sv_clone": (C.c_int, [_P, C.POINTER(_P)]),
    "qcsim_sv_sync": (C.c_int, [_P]),
    "qcsim_sv_flush": (C.c_int, [_P]),
    "qcsim_sv_n_qubits": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "qcsim_sv_device_ptr": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "qcsim_sv_set_basis_state": (C.c_int, [_P, _U64]),
    "qcsim_sv_fill": (C.c_int, [_P, C.c_double, C.c_double]),
    "qcsim_sv_set_amplitude": (C.c_int, [_P, _U64, C.c_double, C.c_double]),
    "qcsim_sv_get_amplitude": (C.c_int, [_P, _U64, _DP]),
    "qcsim_sv_upload": (C.c_int, [_P, C.c_void_p, _U64, _U64]),
    "qcsim_sv_download": (C.c_int, [_P, C.c_void_p, _U64, _U64]),
    "qcsim_sv_norm2": (C.c_int, [_P, _DP]),
    "qcsim_sv_scale": (C.c_int, [_P, C.c_double]),
    "qcsim_sv_normalize": (C.c_int, [_P]),
    "qcsim_sv_save_state": (C.c_int, [_P]),
    "qcsim_sv_restore_state": (C.c_int, [_P, C.c_int]),
    "qcsim_sv_inner_product": (C.c_int, [_P, _P, _DP]),
    "qcsim_sv_apply": (C.c_int, [_P, C.c_int, C.c_void_p, C.c_int, _U64, _U64, _U64]),
    "qcsim_sv_apply_batch": (C.c_int, [_P, C.c_void_p, _U64]),
    "qcsim_sv_set_fusion": (C.c_int, [_P, C.c_int]),
    "qcsim_sv_apply_operator": (C.c_int, [_P, C.c_void_p]),
    "qcsim_circuit_save": (C.c_int, [C.c_char_p, C.c_uint32, C.c_void_p, _U64]),
    "qcsim_circuit_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), _U64P]),
    "qcsim_circuit_free": (None, [C.c_void_p]),
    "qcsim_sv_apply_circuit_file": (C.c_int, [_P, C.c_char_p]),
    "qcsim_sv_qft": (C.c_int, [_P, _U64, _U64, C.c_int, C.c_int]),
    "qcsim_sv_measure_all": (C.c_int, [_P, C.c_double, _U64P]),
    "qcsim_sv_measure": (C.c_int, [_P, _U64, _U64, C.c_double, _U64P]),
    "qcsim_sv_measure_all_nocollapse": (C.c_int, [_P, C.c_double, _U64P]),
    "qcsim_sv_measure_nocollapse": (C.c_int, [_P, _U64, _U64, C.c_double, _U64P]),
    "qcsim_sv_qubit_probability": (C.c_int, [_P, _U64, _DP]),
    "qcsim_sv_sample": (C.c_int, [_P, C.c_void_p, _U64, C.c_void_p]),
    "qcsim_sv_set_strict_measure": (C.c_int, [_P, C.c_int]),
    "qcsim_sv_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "qcsim_sv_reset_stats": (C.c_int, [_P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA extension; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m qcsim_b200.build` (needs nvcc). "
            "qcsim_b200 has no CPU fallback.")
    try:
        # libqcsim_b200.so links libnccl.so.2; torch bundles a newer build under the same soname.
        # Whichever is loaded first serves the whole process, and libtorch_cuda needs its own one:
        # let torch load it first (torch is plumbing here: device memory, streams, torch.distributed).
        import torch  # noqa: F401
    except ImportError:
        pass
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI mismatch, which must be loud
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        msg = load().qcsim_last_error()
        raise QcsimError(rc, msg.decode() if msg else "")
