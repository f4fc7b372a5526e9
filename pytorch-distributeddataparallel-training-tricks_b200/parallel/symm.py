"""NVLink symmetric memory + the fused all-reduce back end of the DDP reducer.

Symmetric memory = one allocation per rank, each mapped into every rank's address space (and, on
NVSwitch systems that support it, into one multicast address).  Two providers:

``torch``  ``torch.distributed._symmetric_memory`` used purely as allocator + rendezvous (CUDA VMM
           handles exchanged through the process group's store, multicast bind) -- SURVEY.md 5.8 step 1.
           All kernels operating on the memory are ours.
``ipc``    fallback: ``cudaMalloc`` + CUDA IPC handles exchanged with ``all_gather_object``.  Peer
           loads/stores work, multicast does not.

:class:`FusedNvlinkBackend` implements the reducer back-end protocol (``launch`` / ``finish`` plus the
stream-explicit ``launch_on`` / ``join`` used under CUDA-graph capture).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from ..ops import require_native

ALGO = {"auto": 0, "oneshot": 1, "twoshot": 2, "multimem": 3}


class SymmetricRegion:
    """A float32 region + a uint32 signal pad present on every rank, with peer pointers."""

    def __init__(self, numel: int, device: torch.device, group=None, provider: str = "auto"):
        self.ext = require_native()
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.numel = int(numel)
        self.provider = None
        self.mc_ptr = 0
        self._keep = []
        sig_words = self.ext.signal_pad_bytes() // 4
        err = None
        want = os.environ.get("DDPT_SYMM_PROVIDER", provider)
        if want in ("auto", "torch"):
            try:
                self._init_torch(sig_words)
                self.provider = "torch"
            except Exception as exc:  # driver / fabric / seccomp restrictions
                err = exc
                if want == "torch":
                    raise
        if self.provider is None:
            try:
                self._init_ipc(sig_words)
                self.provider = "ipc"
            except Exception as exc:
                raise RuntimeError(f"no symmetric-memory provider works (torch: {err!r}; ipc: {exc!r})")
        # agree on the provider (a mixed world would deadlock later)
        flags = [None] * self.world
        dist.all_gather_object(flags, self.provider, group=self.group)
        if len(set(flags)) != 1:
            raise RuntimeError(f"ranks disagree on the symmetric-memory provider: {flags}")
        dist.barrier(group=self.group)

    def _init_torch(self, sig_words: int) -> None:
        import torch.distributed._symmetric_memory as symm_mem
        gname = self.group.group_name
        buf = symm_mem.empty(self.numel, dtype=torch.float32, device=self.device)
        sig = symm_mem.empty(sig_words, dtype=torch.int32, device=self.device)
        hb = symm_mem.rendezvous(buf, gname)
        hs = symm_mem.rendezvous(sig, gname)
        buf.zero_(); sig.zero_()
        self.tensor, self.signal = buf, sig
        self.peer_ptrs = [int(p) for p in hb.buffer_ptrs]
        self.signal_ptrs = [int(p) for p in hs.buffer_ptrs]
        try:
            self.mc_ptr = int(hb.multicast_ptr) if os.environ.get("DDPT_NO_MULTICAST") != "1" else 0
        except Exception:
            self.mc_ptr = 0
        self._keep += [hb, hs]

    def _init_ipc(self, sig_words: int) -> None:
        ext, dev = self.ext, self.device.index if self.device.index is not None else torch.cuda.current_device()
        bptr, bh = ext.ipc_alloc(self.numel * 4, dev)
        sptr, sh = ext.ipc_alloc(sig_words * 4, dev)
        handles: List[Optional[tuple]] = [None] * self.world
        dist.all_gather_object(handles, (bytes(bh), bytes(sh), os.getpid()), group=self.group)
        self.peer_ptrs, self.signal_ptrs = [], []
        for r, (hb, hs, pid) in enumerate(handles):
            if r == self.rank:
                self.peer_ptrs.append(bptr); self.signal_ptrs.append(sptr)
            else:
                self.peer_ptrs.append(ext.ipc_open(hb, dev)); self.signal_ptrs.append(ext.ipc_open(hs, dev))
        self.tensor = ext.tensor_from_ptr(bptr, self.numel, dev, False)
        self.signal = ext.tensor_from_ptr(sptr, sig_words, dev, True)


class FusedNvlinkBackend:
    """Reducer back end: one fused kernel per bucket (reduce + 1/(world*scale) + overflow flag)."""

    fuses_unscale = True

    def __init__(self, device: torch.device, group=None, algo: str = "auto"):
        self.ext = require_native()
        self.device = torch.device(device)
        self.group = group
        self.algo = ALGO[os.environ.get("DDPT_ALLREDUCE_ALGO", algo)]
        self.region: Optional[SymmetricRegion] = None
        self._stream = torch.cuda.Stream(device=self.device)
        self._launched = False
        self._events: List[torch.cuda.Event] = []
        self._unit = torch.tensor([1.0, 1.0, 0.0, 0.0, 0.0], dtype=torch.float32, device=self.device)
        self.launches = 0

    # -- allocation: the DDP gradient arena lives in symmetric memory --------------------------------
    def allocate(self, numel: int) -> torch.Tensor:
        self.region = SymmetricRegion(numel, self.device, self.group)
        return self.region.tensor

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
        self.ext.fused_allreduce(r.peer_ptrs, r.signal_ptrs, r.mc_ptr, r.rank, bucket.start, bucket.numel, state,
                                 state[2:3], bucket.index % 8, self.algo, 0)
        self.launches += 1

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

    def broadcast(self, offset: int, numel: int, root: int = 0) -> None:
        r = self.region
        self.ext.symm_broadcast(r.peer_ptrs, r.signal_ptrs, r.rank, root, offset, numel, 7)
