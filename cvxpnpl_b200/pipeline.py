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


def part_bounds(B, fracs):
    """Contiguous sub-batch boundaries [0, ..., B] for the fractions `fracs` (normalised, monotone, exact at both ends)."""
    tot = float(sum(fracs))
    acc, bounds = 0.0, [0]
    for f in list(fracs)[:-1]:
        acc += float(f) / tot
        bounds.append(max(bounds[-1], min(B, int(round(acc * B)))))
    bounds.append(B)
    return bounds


class HostStager:
    """One batch from pinned HOST memory with the copy hidden as far as a single step allows.

    parts == 1: the correspondences go up in `chunks` slices on a copy stream and the pre-pass
    kernels (assembly + start decomposition + early iterations, `cvxpnpl_b200_prepass`) of each
    slice run as soon as that slice has landed, i.e. under the copies of the following slices.
    Everything after the pre-pass needs the whole batch and follows on the current stream: the
    solver only starts when the last byte has arrived (64 MB, ~1.3 ms for 1e5 PnPL problems),
    and the record only leaves when the last problem is done.

    parts > 1: the problems are independent, so the batch is solved as `parts` contiguous
    sub-batches, each a complete solve on a stream of its own with its own workspace: sub-batch
    p starts when ITS slice has arrived, the slices of the later ones arrive under its kernels,
    and its `[n, 15]` rows go back to the host (`host_record`) under the kernels of the next
    ones.  Persistent kernels of consecutive sub-batches hand the SMs over CTA by CTA, so the
    GPU stays as busy as with one solve; only the first slice's copy and the last sub-batch's
    rows are outside the compute."""

    # problems in the LAST slice: its pre-pass is the only one that is not hidden behind a copy, and the pre-pass
    # kernels need the same time for anything up to one wave (148 SMs x 2 CTAs x 64 problems)
    LAST_SLICE = 12288

    def __init__(self, K, device, chunks=4, parts=1, **solve_kwargs):
        self.device = torch.device(device)
        self.K = torch.as_tensor(K, dtype=torch.float64).to(self.device)
        self.chunks = int(chunks)
        # parts: a count (equal sub-batches) or a sequence of fractions of the batch, e.g. (0.2, 0.4, 0.4): a short
        # first sub-batch lets the kernels start sooner
        if isinstance(parts, (tuple, list)):
            tot = float(sum(parts))
            self.part_fracs = [float(f) / tot for f in parts]
            self.parts = len(self.part_fracs)
        else:
            self.parts = max(1, int(parts))
            self.part_fracs = [1.0 / self.parts] * self.parts
        self.kw = solve_kwargs
        self.copy_stream = torch.cuda.Stream(self.device)
        self.buf = None
        self.ws = None
        self.out = None
        self.done = None     # event: the previous solve no longer reads the input buffers
        self.part_streams = None
        self.part_ws = None
        self.part_out = None

    def _solve_parts(self, host, B, record, host_record) -> BatchedPoses:
        dev, P = self.device, self.parts
        bounds = part_bounds(B, self.part_fracs)
        if self.buf is None or self.buf.keys() != host.keys() or any(self.buf[k].shape != v.shape for k, v in host.items()):
            self.buf = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
            self.part_ws = self.part_out = None
        if self.part_streams is None:
            self.part_streams = [torch.cuda.Stream(dev) for _ in range(P)]
        if self.part_ws is None:
            self.part_ws = [Workspace(hi - lo, dev) for lo, hi in zip(bounds[:-1], bounds[1:])]
            with torch.cuda.device(dev):
                self.out = BatchedPoses(
                    R=torch.empty((B, 4, 3, 3), dtype=torch.float64, device=dev),
                    t=torch.empty((B, 4, 3), dtype=torch.float64, device=dev),
                    n_poses=torch.empty(B, dtype=torch.int32, device=dev),
                    status=torch.empty(B, dtype=torch.int32, device=dev),
                    iters=torch.empty(B, dtype=torch.int32, device=dev),
                    obj=torch.empty((B, 2), dtype=torch.float64, device=dev))
                self.own_record = torch.empty((B, RECORD), dtype=torch.float64, device=dev)
            o = self.out
            self.part_out = [BatchedPoses(R=o.R[lo:hi], t=o.t[lo:hi], n_poses=o.n_poses[lo:hi], status=o.status[lo:hi],
                                          iters=o.iters[lo:hi], obj=o.obj[lo:hi]) for lo, hi in zip(bounds[:-1], bounds[1:])]
        if record is None:
            record = self.own_record
        main = torch.cuda.current_stream(dev)
        entry = torch.cuda.Event()
        entry.record(main)
        events = []
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
        launches = 0
        for p, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
            if hi <= lo:
                continue
            sp = self.part_streams[p]
            with torch.cuda.stream(sp):
                sp.wait_event(entry)
                sp.wait_event(events[p])
                sub = {k: v[lo:hi] for k, v in self.buf.items()}
                po = solve_batched(self.K, pts_2d=sub.get("pts_2d"), pts_3d=sub.get("pts_3d"), line_2d=sub.get("line_2d"),
                                   line_3d=sub.get("line_3d"), workspace=self.part_ws[p], out=self.part_out[p],
                                   record=record[lo:hi], **self.kw)
                launches += po.launches
                if host_record is not None:
                    host_record[lo:hi].copy_(record[lo:hi], non_blocking=True)   # under the kernels of the next parts
            main.wait_stream(sp)
        self.out.launches = launches
        self.out.record = record
        self.done = torch.cuda.Event()
        self.done.record(main)
        return self.out

    def solve(self, host: Dict[str, torch.Tensor], record=None, host_record=None) -> BatchedPoses:
        """`record`: optional [B,15] CUDA tensor for the packed rows (see solve_batched); `host_record`: optional
        pinned [B,15] host tensor that receives those rows (valid after a synchronisation of the current stream)."""
        host = {k: v for k, v in host.items() if v is not None and v.shape[1] > 0}
        B = next(iter(host.values())).shape[0]
        if self.parts > 1 and B >= self.parts * 4096:
            return self._solve_parts(host, B, record, host_record)
        if (self.buf is None or self.ws is None or self.buf.keys() != host.keys()
                or any(self.buf[k].shape != v.shape for k, v in host.items())):
            self.buf = {k: torch.empty_like(v, device=self.device) for k, v in host.items()}
            with torch.cuda.device(self.device):       # the workspace is sized for THIS device's SM count
                self.ws, self.out = Workspace(B, self.device), None
            self.part_ws = self.part_out = None
        if host_record is not None and record is None:
            record = torch.empty((B, RECORD), dtype=torch.float64, device=self.device)
        main = torch.cuda.current_stream(self.device)
        if self.chunks > 1 and B > self.chunks * self.LAST_SLICE:
            head = B - self.LAST_SLICE          # equal slices, then a short last one
            bounds = [(c * head) // (self.chunks - 1) for c in range(self.chunks)] + [B]
        else:
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
                    n += int(lib.cvxpnpl_b200_last_launch_count())
            return n

        self.out = solve_batched(self.K, pts_2d=self.buf.get("pts_2d"), pts_3d=self.buf.get("pts_3d"),
                                 line_2d=self.buf.get("line_2d"), line_3d=self.buf.get("line_3d"), workspace=self.ws,
                                 out=self.out, _prepass_hook=hook, record=record, **self.kw)
        if host_record is not None:
            host_record.copy_(record, non_blocking=True)
        self.done = torch.cuda.Event()
        self.done.record(main)
        return self.out
