"""NVLink symmetric memory + the fused all-reduce back end of the DDP reducer.

Symmetric memory = one allocation per rank, each mapped into every rank's address space (and, on
NVSwitch systems that support it, into one multicast address).  Two providers:

``torch``  ``torch.distributed._symmetric_memory`` used purely as allocator + rendezvous (CUDA VMM
           handles exchanged through the process group's store, multicast bind) -- SURVEY.md 5.8 step 1.
           All kernels operating on the memory are ours.
``ipc``    fallback: ``cudaMalloc`` + CUDA IPC handles exchanged with ``all_gather_object``.  Peer
           loads/stores work, multicast does not.

A :class:`SymmetricRegion` is what one DDP instance owns: the fp32 gradient arena, the signal pad of the
flag barriers, and a scratch region (push slots of the one-shot all-reduce + the 16-bit staging of the bf16
wire format).  :class:`FusedNvlinkBackend` implements the reducer back-end protocol (``launch`` /
``finish`` plus the stream-explicit ``launch_on`` / ``join`` used under CUDA-graph capture) and the
rank-0 broadcast of the module's buffers (BatchNorm statistics; reference: DDP ``_sync_buffers`` behind
utils/process.py:20).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from ..ops import require_native

ALGO = {"auto": 0, "oneshot": 1, "twoshot": 2, "multimem": 3}
WIRE = {"fp32": 0, "bf16": 1}


class SymmBuffer:
    """One symmetric allocation: local tensor, every rank's mapping of it, optional multicast mapping."""

    def __init__(self, tensor: torch.Tensor, peer_ptrs: List[int], mc_ptr: int, keep=None):
        self.tensor, self.peer_ptrs, self.mc_ptr, self._keep = tensor, peer_ptrs, mc_ptr, keep


class SymmetricAllocator:
    def __init__(self, device: torch.device, group=None, provider: str = "auto"):
        self.ext = require_native()
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.provider: Optional[str] = None
        self._want = os.environ.get("DDPT_SYMM_PROVIDER", provider)
        self._torch_error: Optional[BaseException] = None

    # -- providers -------------------------------------------------------------------------------------
    def _alloc_torch(self, numel: int, dtype: torch.dtype) -> SymmBuffer:
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(numel, dtype=dtype, device=self.device)
        handle = symm_mem.rendezvous(buf, self.group.group_name)
        buf.zero_()
        mc = 0
        if os.environ.get("DDPT_NO_MULTICAST") != "1":
            try:
                mc = int(handle.multicast_ptr)
            except Exception:
                mc = 0
        return SymmBuffer(buf, [int(p) for p in handle.buffer_ptrs], mc, keep=handle)

    def _alloc_ipc(self, numel: int, dtype: torch.dtype) -> SymmBuffer:
        ext = self.ext
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        itemsize = torch.empty((), dtype=dtype).element_size()
        ptr, handle = ext.ipc_alloc(numel * itemsize, dev)
        handles: List[Optional[tuple]] = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        peers = [ptr if r == self.rank else ext.ipc_open(h, dev) for r, h in enumerate(handles)]
        words = numel * itemsize // 4
        raw = ext.tensor_from_ptr(ptr, words, dev, dtype != torch.float32)
        tensor = raw if dtype in (torch.float32, torch.int32) else raw.view(dtype)
        return SymmBuffer(tensor, peers, 0)

    def alloc(self, numel: int, dtype: torch.dtype = torch.float32) -> SymmBuffer:
        if self.provider is None:
            self._pick_provider(numel, dtype)
            return self._first
        return self._alloc_torch(numel, dtype) if self.provider == "torch" else self._alloc_ipc(numel, dtype)

    def _pick_provider(self, numel: int, dtype: torch.dtype) -> None:
        mine, first = None, None
        if self._want in ("auto", "torch"):
            try:
                first = self._alloc_torch(numel, dtype)
                mine = "torch"
            except Exception as exc:  # driver / fabric / seccomp restrictions
                self._torch_error = exc
                if self._want == "torch":
                    raise
        # agree on the provider before anyone touches the fallback (a mixed world would deadlock later)
        flags = [None] * self.world
        dist.all_gather_object(flags, mine, group=self.group)
        if all(f == "torch" for f in flags):
            self.provider, self._first = "torch", first
            return
        if self._want == "torch":
            raise RuntimeError(f"torch symmetric memory unavailable on some ranks: {flags}")
        try:
            self.provider = "ipc"
            self._first = self._alloc_ipc(numel, dtype)
        except Exception as exc:
            raise RuntimeError(f"no symmetric-memory provider works (torch: {self._torch_error!r}; ipc: {exc!r})")


class SymmetricRegion:
    """fp32 region + signal pad + scratch present on every rank, with peer (and multicast) pointers."""

    def __init__(self, numel: int, device: torch.device, group=None, provider: str = "auto",
                 allocator: Optional[SymmetricAllocator] = None):
        self.alloc = allocator if allocator is not None else SymmetricAllocator(device, group, provider)
        ext = self.ext = self.alloc.ext
        self.device, self.group = self.alloc.device, self.alloc.group
        self.rank, self.world = self.alloc.rank, self.alloc.world
        self.numel = int(numel)
        self._buf = self.alloc.alloc(self.numel, torch.float32)
        self._sig = self.alloc.alloc(ext.signal_pad_bytes() // 4, torch.int32)
        self._scratch = self.alloc.alloc(ext.scratch_bytes(self.numel), torch.uint8)
        self.provider = self.alloc.provider
        self.tensor, self.signal, self.scratch = self._buf.tensor, self._sig.tensor, self._scratch.tensor
        self.peer_ptrs, self.mc_ptr = self._buf.peer_ptrs, self._buf.mc_ptr
        self.signal_ptrs = self._sig.peer_ptrs
        self.scratch_ptrs, self.mc_scratch_ptr = self._scratch.peer_ptrs, self._scratch.mc_ptr
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)


class FusedNvlinkBackend:
    """Reducer back end: one fused kernel per bucket (cast + reduce + 1/(world*scale) + overflow flag)."""

    fuses_unscale = True
    uses_comm_stream = True

    def __init__(self, device: torch.device, group=None, algo: str = "auto", wire: str = "fp32", blocks: int = 0):
        self.ext = require_native()
        self.device = torch.device(device)
        self.group = group
        self.algo = ALGO[os.environ.get("DDPT_ALLREDUCE_ALGO", algo)]
        self.wire_name = os.environ.get("DDPT_COMM_WIRE", wire)
        self.wire = WIRE[self.wire_name]
        self.blocks = int(os.environ.get("DDPT_COMM_BLOCKS", blocks))
        self.allocator = SymmetricAllocator(self.device, group)
        self.region: Optional[SymmetricRegion] = None
        self.buffers: Optional[SymmBuffer] = None
        self._stream = torch.cuda.Stream(device=self.device)
        self._launched = False
        self._unit = torch.tensor([1.0, 1.0, 0.0, 0.0, 0.0], dtype=torch.float32, device=self.device)
        self.launches = 0
        self.broadcasts = 0

    # -- allocation: the DDP gradient arena lives in symmetric memory --------------------------------
    def allocate(self, numel: int) -> torch.Tensor:
        self.region = SymmetricRegion(numel, self.device, self.group, allocator=self.allocator)
        return self.region.tensor

    def allocate_buffers(self, numel: int) -> torch.Tensor:
        """Symmetric fp32 region that holds the module's buffers (BatchNorm statistics), 16-byte granular."""
        self.buffers = self.allocator.alloc((int(numel) + 3) // 4 * 4, torch.float32)
        return self.buffers.tensor

    def bind(self, arena) -> None:
        assert self.region is not None and arena.grad_arena.data_ptr() == self.region.tensor.data_ptr()

    @property
    def provider(self) -> str:
        return self.region.provider if self.region is not None else "none"

    @property
    def multicast(self) -> bool:
        return self.region is not None and self.region.mc_ptr != 0

    # -- kernel launch --------------------------------------------------------------------------------
    def _kernel(self, bucket, scaler) -> None:
        r = self.region
        state = scaler.state if (scaler is not None and scaler.on_device) else self._unit
        self.ext.fused_allreduce(r.peer_ptrs, r.signal_ptrs, r.mc_ptr, r.scratch_ptrs, r.mc_scratch_ptr, r.rank,
                                 bucket.start, bucket.numel, state, state[2:3], bucket.index % 7, self.algo, self.wire,
                                 self.blocks, 0)
        self.launches += 1       # channels 0..6 carry buckets, channel 7 the buffer broadcast

    def launch_on(self, comm_stream, main_stream, arena, bucket, scaler) -> None:
        """Fork: comm_stream waits for what main_stream has enqueued so far, then runs the bucket kernel."""
        ev = torch.cuda.Event()
        ev.record(main_stream)
        comm_stream.wait_event(ev)
        with torch.cuda.stream(comm_stream):
            self._kernel(bucket, scaler)

    def join(self, comm_stream, main_stream) -> None:
        ev = torch.cuda.Event()
        ev.record(comm_stream)
        main_stream.wait_event(ev)

    # reducer protocol (autograd-hook driven path)
    def launch(self, arena, bucket, scaler=None) -> None:
        self.launch_on(self._stream, torch.cuda.current_stream(), arena, bucket, scaler)
        self._launched = True

    def finish(self) -> None:
        if self._launched:
            self.join(self._stream, torch.cuda.current_stream())
            self._launched = False

    # -- rank-0 -> all broadcast of the buffer region (one launch of our kernel on the current stream) ----
    def broadcast_buffers(self, numel: int, root: int = 0) -> None:
        b, r = self.buffers, self.region
        self.ext.symm_broadcast(b.peer_ptrs, r.signal_ptrs, b.mc_ptr, r.rank, root, 0, (int(numel) + 3) // 4 * 4, 7, 0)
        self.broadcasts += 1
