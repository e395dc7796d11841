"""Peer-memory exchange buffers for the multi-GPU loss path (host side of csrc/exchange.cu).

The reference gathers the features of all ranks with two `torch.distributed` all-gathers inside every
`get_loss` call (cn_clip/training/train.py:53-84).  Here every rank owns ONE exchange buffer that all its
peers map through CUDA IPC; the cast kernel pushes 16-bit rows straight into the peers' buffers over
NVLink and the forward consumes them tile by tile as their arrival flags go up (include/nans_clip.h,
"multi-GPU exchange").  This module only does the plumbing that must happen on the host, once per
process group and buffer size:

  * allocate the buffer in the library (`nans_peer_alloc`: cudaMalloc + IPC handle),
  * exchange the 64-byte handles over the process group (`all_gather_object`),
  * map the peers (`nans_peer_open`) and keep the `nans_xchg_t` descriptor the kernels take.

No collective runs per step.  If IPC mapping is not possible on some rank (different nodes, no peer
access) the group falls back, as a whole, to the NCCL all-gather path of loss.py.
"""
from __future__ import annotations

import ctypes
import os
import warnings
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib

MAX_PEERS = 16
FLAG_ROWS = 64


class XchgDesc(ctypes.Structure):
    """Mirror of nans_xchg_t (include/nans_clip.h)."""
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("base", ctypes.c_void_p * MAX_PEERS),
                ("n_loc", ctypes.c_int64), ("D", ctypes.c_int64),
                ("feat_off", ctypes.c_int64), ("lse_off", ctypes.c_int64), ("lse_len", ctypes.c_int64),
                ("fflag_off", ctypes.c_int64), ("lflag_off", ctypes.c_int64), ("bytes", ctypes.c_int64),
                ("epoch", ctypes.c_void_p)]


def layout_bytes(world: int, n_loc: int, D: int) -> int:
    d = XchgDesc()
    d.world = world
    _lib.check(_lib.load().nans_xchg_layout(ctypes.byref(d), n_loc, D))
    return int(d.bytes)


def make_desc(world: int, rank: int, bases, n_loc: int, D: int, epoch_ptr: int) -> XchgDesc:
    """A descriptor over explicit base pointers (the IPC-mapped peers, or — tests — W buffers of one
    process that emulate the ranks one after the other)."""
    d = XchgDesc()
    d.world, d.rank = world, rank
    _lib.check(_lib.load().nans_xchg_layout(ctypes.byref(d), n_loc, D))
    for r in range(world):
        d.base[r] = int(bases[r])
    d.epoch = int(epoch_ptr)
    return d


def eligible(n_loc: int, D: int, world: int) -> bool:
    """Shapes the push path serves: whole 256-column tiles per rank, the narrow-pair backward."""
    return world > 1 and world <= MAX_PEERS and n_loc > 0 and n_loc % 256 == 0 and D % 8 == 0 and D <= 1024 and \
        os.environ.get("NANS_BWD_NP", "1") != "0"


class PeerExchange:
    """The exchange state of one process group on this rank's device."""

    def __init__(self, group):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.own: Optional[int] = None          # device pointer of this rank's buffer
        self.capacity = 0
        self.bases: list[int] = []
        self.shape: Optional[tuple[int, int]] = None
        self.desc: Optional[XchgDesc] = None
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.dev)   # steps completed (device word)
        self.forwards = 0                        # host mirror of the number of forwards issued
        self.push_stream = torch.cuda.Stream(self.dev)   # the NVLink push runs under the forward
        self.push_stream_b = torch.cuda.Stream(self.dev)  # hybrid push: the stream of the small push kernel
        self.fork = torch.cuda.Event()
        self.fork2 = torch.cuda.Event()
        self.join = torch.cuda.Event()
        self.join_b = torch.cuda.Event()
        self.broken = False

    # ---- allocation / mapping (collective over the group) -------------------------------------
    def _release(self) -> None:
        lib = _lib.load()
        for r, b in enumerate(self.bases):
            if r != self.rank and b:
                lib.nans_peer_close(ctypes.c_void_p(b))
        if self.own:
            lib.nans_peer_free(ctypes.c_void_p(self.own))
        self.own, self.bases, self.capacity = None, [], 0

    def _allocate(self, nbytes: int) -> bool:
        lib = _lib.load()
        ok, handle, ptr = True, ctypes.create_string_buffer(64), ctypes.c_void_p()
        try:
            _lib.check(lib.nans_peer_alloc(nbytes, ctypes.byref(ptr), handle))
        except Exception:  # noqa: BLE001 - the group decides together below
            ok = False
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, handle.raw if ok else b""), group=self.group)
        if not all(h[0] for h in handles):
            if ok:
                lib.nans_peer_free(ptr)
            return False
        bases, opened = [0] * self.world, True
        for r, (_, raw) in enumerate(handles):
            if r == self.rank:
                bases[r] = ptr.value
                continue
            p = ctypes.c_void_p()
            if lib.nans_peer_open(raw, ctypes.byref(p)) != 0:
                opened = False
                break
            bases[r] = p.value
        flag = torch.tensor([1 if opened else 0], device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self.own, self.bases, self.capacity = ptr.value, bases, nbytes
        if int(flag.item()) == 0:
            self._release()
            return False
        return True

    def ensure(self, n_loc: int, D: int) -> Optional[XchgDesc]:
        """The descriptor for (n_loc, D); (re)allocates / re-zeroes collectively when the layout changes.
        Returns None when peer mapping is unavailable (the caller uses the NCCL path)."""
        if self.broken:
            return None
        if self.shape == (n_loc, D):
            return self.desc
        need = layout_bytes(self.world, n_loc, D)
        # every rank's kernels of the old layout must have finished before anything is remapped or zeroed
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)
        if need > self.capacity:
            self._release()
            if not self._allocate(max(need, 64 << 20)):
                self.broken = True
                warnings.warn("nans_clip_b200: CUDA IPC peer mapping is unavailable on this group; the loss uses the "
                              "NCCL all-gather path instead of the NVLink push exchange")
                return None
        else:
            # a new layout moves the flag words: stale bytes there must not read as "arrived"
            _lib.check(_lib.load().nans_peer_zero(ctypes.c_void_p(self.own), self.capacity,
                                                  torch.cuda.current_stream(self.dev).cuda_stream))
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)
        self.shape = (n_loc, D)
        self.desc = make_desc(self.world, self.rank, self.bases, n_loc, D, self.epoch.data_ptr())
        return self.desc


_EXCHANGES: dict = {}


def for_group(group) -> Optional[PeerExchange]:
    """The group's exchange (created on first use), or None when disabled (NANS_EXCHANGE=nccl) or off CUDA."""
    if os.environ.get("NANS_EXCHANGE", "push") != "push" or not torch.cuda.is_available():
        return None
    key = id(group)
    ex = _EXCHANGES.get(key)
    if ex is None:
        ex = _EXCHANGES[key] = PeerExchange(group)
    return ex
