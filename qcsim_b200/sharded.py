"""Sharded registers: one process per GPU, the state split on its top log2(world) qubits.

torch.distributed is plumbing only (it carries the 128-byte NCCL id from rank 0 to the other
ranks); the data path -- global<->local qubit exchanges over NVLink and the scalar all-reduces of
the measurement path -- lives in libqcsim_b200.so (csrc/dist.cu).
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from .register import QubitRegister


class ShardedQubitRegister(QubitRegister):
    """QubitRegister whose amplitudes live on `world` GPUs.  Same API; every rank must make the
    same calls in the same order (SPMD).  getRegisterStorage() returns this rank's slice."""

    def __init__(self, N: int, device: int, rank: int, world: int, nccl_id: bytes, seed: int = 1):
        lib = _lib.load()
        h = C.c_void_p()
        buf = C.create_string_buffer(nccl_id, 128)
        _lib.check(lib.qcsim_sv_create_sharded(C.byref(h), N, device, rank, world, buf))
        super().__init__(N, seed=seed, _handle=h)
        self.rank, self.world = rank, world
        nl = C.c_int()
        _lib.check(lib.qcsim_sv_n_qubits(h, None, C.byref(nl)))
        self.n_local = nl.value
        self.slice_first = rank << self.n_local
        self.slice_count = 1 << self.n_local

    def getRegisterStorage(self):
        return self.download(self.slice_first, self.slice_count)

    def upload_slice(self, vals):
        self.upload(vals, self.slice_first)


def nccl_unique_id() -> bytes:
    lib = _lib.load()
    buf = C.create_string_buffer(128)
    _lib.check(lib.qcsim_nccl_unique_id(buf))
    return buf.raw


def create_register(n: int, device: int, rank: int = 0, world: int = 1, dist=None, seed: int = 1):
    """Single-GPU register, or a shard of a `world`-GPU register when torch.distributed is up."""
    if world == 1:
        return QubitRegister(n, device=device, seed=seed)
    ids = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return ShardedQubitRegister(n, device, rank, world, ids[0], seed=seed)
