"""Peer-memory communicator for the GPUs of one node (include/b200vae.h, "peer-memory exchange").

The reference is single-process; its sharded counterpart here has two kinds of exchange per step:
ten 1 KB BatchNorm-statistics exchanges in the encoder and one gradient all-reduce.  Both are latency bound, so they
run as kernels that read / write the other GPUs' memory directly over NVLink / NVSwitch (csrc/peer.cuh) instead of
NCCL calls plus torch glue.  torch.distributed is only the out-of-band channel for the CUDA IPC handles.

    comm = PeerComm(process_group)            # collective
    grads = comm.alloc_flat(numel)            # collective; .local is a torch view of this rank's buffer
    comm.allgather(x)                         # [world, n] in a few microseconds
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _C


class _RawCuda:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, numel, owner):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 2,
                                         "strides": None}
        self._owner = owner


class PeerBuffer:
    """One symmetric allocation: `ptrs[r]` is rank r's buffer as addressable from this process."""

    def __init__(self, comm, nbytes):
        lib = _C.load()
        self.comm, self.nbytes = comm, nbytes
        own, handle = C.c_void_p(), C.create_string_buffer(_C.PEER_HANDLE_BYTES)
        self.own, self.ptrs, self._mapped = None, [], []
        # every rank takes part in both object exchanges even if its own step failed, so that a failure is agreed on
        # collectively instead of leaving the others waiting
        rc = lib.b200vae_peer_alloc(nbytes, C.byref(own), handle)
        if rc == 0:
            self.own = own.value
        handles = [None] * comm.world
        dist.all_gather_object(handles, handle.raw if rc == 0 else None, group=comm.group)
        ok = all(h is not None for h in handles)
        if ok:
            for r, h in enumerate(handles):
                if r == comm.rank:
                    self.ptrs.append(self.own)
                    continue
                m = C.c_void_p()
                if lib.b200vae_peer_open(h, C.byref(m)) != 0:
                    ok = False
                    break
                self.ptrs.append(m.value)
                self._mapped.append(m.value)
        torch.cuda.synchronize()
        oks = [None] * comm.world
        dist.all_gather_object(oks, ok, group=comm.group)   # also the barrier: every buffer is zero-filled and mapped
        if not all(oks):
            self.close()
            raise _C.B200VaeError(f"CUDA IPC peer mapping failed (per-rank status {oks}, cudaError "
                                  f"{lib.b200vae_last_cuda_error()})")

    def tensor(self, numel=None):
        numel = self.nbytes // 4 if numel is None else numel
        return torch.as_tensor(_RawCuda(self.own, numel, self), device=self.comm.device)

    def ptr_array(self):
        arr = (C.c_void_p * self.comm.world)(*self.ptrs)
        return arr

    def close(self):
        lib = _C.load()
        for m in self._mapped:
            lib.b200vae_peer_close(m)
        self._mapped = []
        if self.own:
            lib.b200vae_peer_free(self.own)
            self.own = None


class PeerComm:
    def __init__(self, group=None, device=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise _C.B200VaeError("PeerComm needs an initialised torch.distributed process group")
        lib = _C.load()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _C.PEER_MAX_WORLD:
            raise _C.B200VaeError(f"PeerComm supports at most {_C.PEER_MAX_WORLD} ranks on one node")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.num_slots, self.max_payload = lib.b200vae_peer_num_slots(), lib.b200vae_peer_max_payload()
        self.exchange = PeerBuffer(self, lib.b200vae_peer_exchange_bytes())
        self.struct = _C.PeerStruct()
        self.struct.world, self.struct.rank = self.world, self.rank
        for r, p in enumerate(self.exchange.ptrs):
            self.struct.buf[r] = p
        self._next_slot = 0
        self._buffers = [self.exchange]

    @property
    def ref(self):
        return C.byref(self.struct)

    def new_slot(self, count=1):
        """Reserve `count` consecutive slots (call in the same order on every rank)."""
        if self._next_slot + count > self.num_slots:
            self._next_slot = 0      # slots may be shared: the protocol only needs every rank to use the same sequence
        s = self._next_slot
        self._next_slot += count
        return s

    def alloc_flat(self, numel):
        buf = PeerBuffer(self, ((numel + 3) // 4 * 4) * 4)
        self._buffers.append(buf)
        return buf

    def allgather(self, x, slot=None):
        """[world, n] fp32 tensor of every rank's x (n <= max_payload)."""
        lib = _C.load()
        x = x.detach().reshape(-1).float().contiguous()
        n = x.numel()
        out = torch.empty(self.world, n, dtype=torch.float32, device=x.device)
        if slot is None:
            slot = self.num_slots - 1
        _C.check(lib.b200vae_peer_allgather(self.ref, slot, x.data_ptr(), n, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "peer_allgather")
        return out

    def barrier(self, slot=None):
        lib = _C.load()
        _C.check(lib.b200vae_peer_allgather(self.ref, self.num_slots - 1 if slot is None else slot, None, 0, None,
                                            torch.cuda.current_stream().cuda_stream), "peer_barrier")

    def check(self):
        """Raise if any exchange on this rank timed out waiting for a peer (synchronises)."""
        out = C.c_int(0)
        torch.cuda.synchronize()
        _C.check(_C.load().b200vae_peer_timed_out(self.ref, C.byref(out)), "peer_timed_out")
        if out.value:
            raise _C.B200VaeError("a peer-memory exchange timed out waiting for another rank")

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for b in self._buffers:
            b.close()
        self._buffers = []


def try_create(group=None):
    """PeerComm if CUDA IPC peer mapping works between ALL ranks of the group, else None.  PeerBuffer agrees on
    failure collectively, so every rank takes the same branch."""
    try:
        return PeerComm(group)
    except _C.B200VaeError as exc:
        if dist.get_rank(group) == 0:
            print(f"[vae_song_b200] peer-memory exchange unavailable, using NCCL: {exc}", flush=True)
        return None
