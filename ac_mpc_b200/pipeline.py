"""Host -> device -> host pipeline for STREAMS of batches (track sweeps, replays, Monte-Carlo evaluation of a controller).

`SpatialMPC.get_control_batch` / `BatchedMPC.solve_host` are synchronous: a call returns when its results are in host
memory, so the 4.9 MB up and 9.8 MB down of a 4096-instance batch (0.3 ms of PCIe) sit next to 0.58 ms of kernels.  A
caller that has the next batch ready before it needs the previous results can overlap them:

    pipe = HostPipeline(mpc, B=4096, fields=[...])            # mpc: BatchedMPC
    t0 = pipe.submit(paths0, None, vmax0)                      # pinned (or plain) host arrays in; returns at once
    t1 = pipe.submit(paths1, None, vmax1)                      # its H2D and kernels overlap batch 0's D2H
    out0 = pipe.wait(t0)                                       # dict of numpy views of the pipeline's pinned buffers,
    ...                                                        # valid until `depth` further submits

Three CUDA streams (copy-in, compute, copy-out) and `depth` slots of device + pinned host buffers; every batch still pays
its own H2D of the inputs and its own D2H of the results -- they just run while another batch computes.  PyTorch is used
for the streams, events and buffers only; the solve is `acmpc_solve_batch_device`.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import _capi


class HostPipeline:
    def __init__(self, mpc, B: int, fields=None, depth: int = 2):
        import torch

        self._torch = torch
        self.mpc, self.B, self.depth = mpc, int(B), int(depth)
        self.H = mpc.H
        spec = _capi.output_spec(self.H)
        self.fields = list(spec) if fields is None else list(fields)
        dev = torch.device("cuda", mpc.device)
        self.dev = dev
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.slots = []
        for _ in range(self.depth):
            packed, views = mpc.alloc_device_outputs(self.B, self.fields)
            h_packed = torch.empty(packed.numel(), dtype=torch.uint8).pin_memory()
            slot = dict(
                d_paths=torch.empty((self.B, self.H, 3), dtype=torch.float64, device=dev),
                d_off=torch.empty(self.B, dtype=torch.float64, device=dev),
                d_vmax=torch.empty(self.B, dtype=torch.float64, device=dev),
                h_paths=torch.empty((self.B, self.H, 3), dtype=torch.float64).pin_memory(),
                h_off=torch.empty(self.B, dtype=torch.float64).pin_memory(),
                h_vmax=torch.empty(self.B, dtype=torch.float64).pin_memory(),
                packed=packed, views=views, h_packed=h_packed,
                h_views={k: v.numpy() for k, v in mpc.unpack(h_packed, self.B, self.H, self.fields).items()},
                in_ready=torch.cuda.Event(), solved=torch.cuda.Event(), out_done=torch.cuda.Event(), used=False)
            self.slots.append(slot)
        self._n = 0

    def _stage(self, slot, stage, arr):
        """A host tensor the copy engine can read asynchronously: the caller's array if it is pinned, else a staged copy
        (made only once the previous upload FROM this slot's staging buffer has left it)."""
        torch = self._torch
        src = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
        if src.is_pinned():
            return src
        if slot["used"]:
            slot["in_ready"].synchronize()
        stage[: src.shape[0]].copy_(src)
        return stage[: src.shape[0]]

    def submit(self, paths, offsets=None, vmax=None, is_localised: bool = False) -> int:
        torch = self._torch
        s = self.slots[self._n % self.depth]
        B = paths.shape[0]
        if B != self.B:
            raise ValueError(f"this pipeline was built for batches of {self.B} instances")
        with torch.cuda.stream(self.s_in):
            if s["used"]:
                self.s_in.wait_event(s["solved"])       # the kernels that read this slot's inputs are done
            s["d_paths"].copy_(self._stage(s, s["h_paths"], paths), non_blocking=True)
            if offsets is not None:
                s["d_off"].copy_(self._stage(s, s["h_off"], offsets), non_blocking=True)
            if vmax is not None:
                s["d_vmax"].copy_(self._stage(s, s["h_vmax"], vmax), non_blocking=True)
            s["in_ready"].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(s["in_ready"])
            if s["used"]:
                self.s_run.wait_event(s["out_done"])    # the previous results of this slot have left the device
            self.mpc.solve_device(s["d_paths"], s["d_off"] if offsets is not None else None,
                                  s["d_vmax"] if vmax is not None else None, is_localised, out=s["views"], stream=self.s_run)
            s["solved"].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(s["solved"])
            s["h_packed"].copy_(s["packed"], non_blocking=True)     # ONE D2H of the packed results
            s["out_done"].record(self.s_out)
        s["used"] = True
        self._n += 1
        return self._n - 1

    def wait(self, ticket: int) -> Dict[str, np.ndarray]:
        if ticket < self._n - self.depth or ticket >= self._n:
            raise ValueError("ticket is not in flight any more (results are valid for `depth` submits)")
        s = self.slots[ticket % self.depth]
        s["out_done"].synchronize()
        return s["h_views"]

    def bytes_per_batch(self):
        """(h2d, d2h) bytes one batch moves (inputs: paths + v_max; outputs: the packed fields incl. slab padding)."""
        return self.B * (3 * self.H + 1) * 8, int(self.slots[0]["packed"].numel())
