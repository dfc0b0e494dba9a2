"""Streaming use of the batched solver from HOST memory.

`HostPipeline` double-buffers the correspondences of consecutive batches: while the
solver works on batch k (main stream), the pinned-host -> device copy of batch k+1
runs on a copy stream, so in steady state the PCIe transfer (64 MB per 1e5 PnPL
problems, ~1.3 ms) hides behind the ~9.5 ms of compute.  The result comes back as
the packed `[B, 15]` pose record of `cvxpnpl_b200.distributed` in a pinned host
tensor.  Plumbing only (torch streams / events); the arithmetic is
`solve_batched`.
"""
import ctypes
from typing import Dict, Optional

import torch

from . import _lib
from .batched import BatchedPoses, Workspace, solve_batched
from .distributed import RECORD

_KEYS = ("pts_2d", "pts_3d", "line_2d", "line_3d")


class HostPipeline:
    def __init__(self, K, device, **solve_kwargs):
        self.device = torch.device(device)
        self.K = torch.as_tensor(K, dtype=torch.float64).to(self.device)
        self.kw = solve_kwargs
        self.copy_stream = torch.cuda.Stream(self.device)
        self.buf = [None, None]          # device input buffers (dict per slot)
        self.staged = [None, None]       # event: H2D into slot finished
        self.free = [None, None]         # event: solve reading slot finished
        self.cur = 0
        self.ws: Optional[Workspace] = None
        self.out: Optional[BatchedPoses] = None
        self.host_out: Optional[torch.Tensor] = None

    def _stage(self, slot: int, host: Dict[str, torch.Tensor]):
        """Enqueue the H2D copy of one batch into `slot` on the copy stream."""
        with torch.cuda.stream(self.copy_stream):
            if self.free[slot] is not None:
                self.copy_stream.wait_event(self.free[slot])      # the previous solve on this slot is done
            live = {k: v for k, v in host.items() if v is not None}
            cur = self.buf[slot]
            if cur is None or cur.keys() != live.keys() or any(cur[k].shape != v.shape for k, v in live.items()):
                # first use of the slot, or the batch changed shape: (re)allocate the staged buffers
                self.buf[slot] = {k: torch.empty_like(v, device=self.device) for k, v in live.items()}
            for k, v in host.items():
                if v is not None:
                    self.buf[slot][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self.staged[slot] = ev

    def step(self, host: Dict[str, torch.Tensor], host_next: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """Solve the batch `host` (pinned host tensors, keys pts_2d / pts_3d / line_2d /
        line_3d; staged by the previous call if it was passed as `host_next`), start the copy of
        `host_next`, and return the pinned `[B, 15]` record (valid after a stream sync)."""
        main = torch.cuda.current_stream(self.device)
        slot = self.cur
        if self.staged[slot] is None:
            self._stage(slot, host)
        main.wait_event(self.staged[slot])
        self.staged[slot] = None
        if host_next is not None:
            self._stage(1 - slot, host_next)                      # overlaps the solve below
        inp = self.buf[slot]
        B = next(iter(inp.values())).shape[0]
        if self.ws is None or self.ws_batch != B:
            with torch.cuda.device(self.device):       # the workspace is sized for THIS device's SM count
                self.ws, self.ws_batch, self.out = Workspace(B, self.device), B, None
            self.rec = torch.empty((B, RECORD), dtype=torch.float64, device=self.device)
            self.host_out = torch.empty((B, RECORD), dtype=torch.float64).pin_memory()
        # the finish kernel writes the packed [B,15] rows itself (desc.record): no packing pass
        self.out = solve_batched(self.K, pts_2d=inp.get("pts_2d"), pts_3d=inp.get("pts_3d"), line_2d=inp.get("line_2d"),
                                 line_3d=inp.get("line_3d"), workspace=self.ws, out=self.out, record=self.rec, **self.kw)
        ev = torch.cuda.Event()
        ev.record(main)
        self.free[slot] = ev
        self.host_out.copy_(self.rec, non_blocking=True)
        self.cur = 1 - slot
        return self.host_out


class HostStager:
    """One batch from pinned HOST memory with the copy hidden as far as a single step allows:
    the correspondences go up in `chunks` slices on a copy stream and the pre-pass kernel
    (assembly + start decomposition, `cvxpnpl_b200_prepass`) of each slice runs as soon as
    that slice has landed, i.e. under the copies of the following slices.  Everything after
    the pre-pass needs the whole batch and follows on the current stream."""

    def __init__(self, K, device, chunks=4, **solve_kwargs):
        self.device = torch.device(device)
        self.K = torch.as_tensor(K, dtype=torch.float64).to(self.device)
        self.chunks = int(chunks)
        self.kw = solve_kwargs
        self.copy_stream = torch.cuda.Stream(self.device)
        self.buf = None
        self.ws = None
        self.out = None
        self.done = None     # event: the previous solve no longer reads the input buffers

    def solve(self, host: Dict[str, torch.Tensor], record=None) -> BatchedPoses:
        """`record`: optional [B,15] CUDA tensor for the packed rows (see solve_batched)."""
        host = {k: v for k, v in host.items() if v is not None and v.shape[1] > 0}
        B = next(iter(host.values())).shape[0]
        if self.buf is None or self.buf.keys() != host.keys() or any(self.buf[k].shape != v.shape for k, v in host.items()):
            self.buf = {k: torch.empty_like(v, device=self.device) for k, v in host.items()}
            with torch.cuda.device(self.device):       # the workspace is sized for THIS device's SM count
                self.ws, self.out = Workspace(B, self.device), None
        main = torch.cuda.current_stream(self.device)
        bounds = [(c * B) // self.chunks for c in range(self.chunks + 1)]
        events = []
        entry = torch.cuda.Event()
        entry.record(main)       # the copies start after everything already queued on the caller's stream
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(entry)
            if self.done is not None:
                self.copy_stream.wait_event(self.done)
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                for k, v in host.items():
                    self.buf[k][lo:hi].copy_(v[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                events.append(ev)

        def hook(lib, d, stream):
            n = 0
            for ev, lo, hi in zip(events, bounds[:-1], bounds[1:]):
                main.wait_event(ev)
                if hi > lo:
                    _lib.check(lib.cvxpnpl_b200_prepass(ctypes.byref(d), lo, hi - lo, ctypes.c_void_p(stream)))
                    n += 1
            return n

        self.out = solve_batched(self.K, pts_2d=self.buf.get("pts_2d"), pts_3d=self.buf.get("pts_3d"),
                                 line_2d=self.buf.get("line_2d"), line_3d=self.buf.get("line_3d"), workspace=self.ws,
                                 out=self.out, _prepass_hook=hook, record=record, **self.kw)
        self.done = torch.cuda.Event()
        self.done.record(main)
        return self.out
