"""Product-level multi-GPU call: one process per GPU, a global batch of independent MPC
instances sharded contiguously over the ranks, NO data-path collective, and ONE exchange at the end that lands every
rank's packed results on rank `dst` (where the caller -- the agent loop -- lives).

Transports of that one exchange (`ShardedMPC(transport=...)`, "auto" tries them in this order):

  "peer"  compute + exchange in ONE kernel: the control / speed kernels of rank r store their outputs STRAIGHT INTO rank
          dst's slab over NVLink -- the output pointers handed to acmpc_solve_batch_device are peer-mapped views of
          dst's buffer (symmetric memory: CUDA VMM handles exchanged once at start-up).  9.8 MB per rank and step at 4096
          instances = 16 GB/s per rank, nothing for NVLink 5.  No collective at all in steady state: the last CTA of the
          control kernel writes a completion flag into dst's memory after a system-scope fence (acmpc_attach_completion),
          dst's consumer stream waits on those flags with cuStreamWaitValue32 (no SM, no kernel -- an NCCL kernel could
          not be scheduled reliably next to the persistent control kernel, which fills every SM's registers, shared and
          tensor memory: measured at 8 GPUs, a per-step NCCL signal cost 0.3 ms of a 0.6 ms step), and dst's own
          launch hands consumed buffers back by writing a credit word into every producer's memory.
  "nccl"  every rank solves into its own HBM, then ONE NCCL gather of the packed buffers to dst (grouped send/recv
          over NVLink).  The round-1 path; also what the CPU tests run over gloo.

Steps are triple-buffered.  "nccl": the gather of step i runs on a side stream while the kernels of step i + 1 run; the
kernels of step i wait for the gather of step i - 2 only (never for the one still in flight).  "peer": the kernels of
step i start only once a LOCAL credit word says dst has released step i - 3, the previous user of the buffer they
write.  Contract of wait() in both: the views are valid until dst's NEXT submit -- whatever reads them must be ordered
before that submit on dst's current stream.

    sh = ShardedMPC(cfg, fields=[...])                       # after dist.init_process_group
    out = sh.solve(paths, offsets, vmax, is_localised)       # HOST global batch in (every rank passes the same arrays,
                                                             # or only its own shard with local=True);
                                                             # dict of numpy arrays for the WHOLE batch on dst, None elsewhere
    t = sh.submit_device(d_paths, d_offsets, d_vmax, loc)    # device-resident shard, asynchronous
    views = sh.wait(t)                                       # dst: dict name -> list of per-rank CUDA views (instance order),
                                                             # valid until the next submit
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from . import _capi
from .sharding import packed_layout, shard_range


NBUF = 3


class ShmUnavailable(RuntimeError):
    """The shared host segment of ShardedMPC's host delivery cannot be created (every rank raises it together)."""


class _Ticket:
    __slots__ = ("buf", "sizes", "B_total", "step")

    def __init__(self, buf, sizes, B_total, step):
        self.buf, self.sizes, self.B_total, self.step = buf, sizes, B_total, step


class ShardedMPC:
    def __init__(self, cfg: "_capi.Config", fields=None, group=None, dst: int = 0, device: Optional[int] = None,
                 transport: str = "auto", solver=None):
        import torch
        import torch.distributed as dist

        self._torch, self._dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dst = int(dst)
        self.H = int(cfg.horizon)
        spec = _capi.output_spec(self.H)
        self.fields = list(spec) if fields is None else list(fields)
        if solver is None:
            from .solver import BatchedMPC

            solver = BatchedMPC(cfg, torch.cuda.current_device() if device is None else device)
        self.solver = solver
        self.cuda = bool(getattr(solver, "device", None) is not None and torch.cuda.is_available())
        self.device = torch.device("cuda", solver.device) if self.cuda else torch.device("cpu")
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError("transport must be auto | peer | nccl")
        self._want = transport
        self.transport = None            # decided at the first allocation
        self.transport_note = ""
        self._cap_B = 0                  # instances per rank the buffers hold
        self._cap = 0                    # bytes per slab
        self._step = 0
        self._side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._signal = None
        self._sig_done = [None] * NBUF
        self._local = [None] * NBUF      # packed output buffer of this rank (a view into dst's slab for "peer")
        self._gathered = [None] * NBUF   # dst: world * cap bytes
        self._symm = None
        self._host_in = None
        self._host_out = None
        self._host_sizes = None
        self._shm = None
        self._shm_failed = False
        self._want_deliver = "auto"
        self._lv_cache, self._gv_cache, self._size_cache = {}, {}, {}
        self._last_B = 0

    # -- buffers --------------------------------------------------------------------------------------------------
    def _ensure(self, B_cap: int):
        """(Re)allocate for shards of up to B_cap instances.  Collective: every rank calls it with the same value."""
        torch, dist = self._torch, self._dist
        if B_cap <= self._cap_B:
            return
        if self.cuda:
            torch.cuda.synchronize(self.device)
        self._cap_B = int(B_cap)
        self._lv_cache, self._gv_cache, self._size_cache = {}, {}, {}
        self._sig_done = [None] * NBUF
        self._step = 0
        self._cap = packed_layout(self._cap_B, self.H, self.fields)[1]
        total = self.world * self._cap
        self.transport = None
        if self.world == 1:
            self.transport = "local"
            for b in range(NBUF):
                self._gathered[b] = torch.empty(total, dtype=torch.uint8, device=self.device)
                self._local[b] = self._gathered[b]
            return
        if self.cuda and self._want in ("auto", "peer"):
            try:
                self._alloc_peer(total)
                self.transport = "peer"
            except Exception as e:      # noqa: BLE001 -- any failure of the VMM / handle exchange means: use NCCL
                self.transport_note = f"peer transport unavailable ({type(e).__name__}: {e})"
                if self._want == "peer":
                    raise
            # every rank must have made the same choice
            ok = torch.tensor([1 if self.transport == "peer" else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self.transport = None
        if self.transport is None:
            self.transport = "nccl"
            for b in range(NBUF):
                self._local[b] = torch.empty(self._cap, dtype=torch.uint8, device=self.device)
                self._gathered[b] = (torch.empty(total, dtype=torch.uint8, device=self.device)
                                     if self.rank == self.dst else None)
        if self.cuda:
            self._signal = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _alloc_peer(self, total: int):
        """Symmetric memory: every rank allocates NBUF x total bytes, the handles are exchanged once, and rank r's output
        buffer becomes the view [r * cap, (r + 1) * cap) of DST's allocation, mapped into r's address space."""
        torch, dist = self._torch, self._dist
        import torch.distributed._symmetric_memory as symm_mem

        grp = self.group if self.group is not None else dist.group.WORLD
        # [ NBUF x world slabs | control: world flag words (128 B apart), then this rank's credit word ]
        data = NBUF * total
        ctrl = (self.world + 1) * 128
        size = data + ctrl
        t = symm_mem.empty(size, dtype=torch.uint8, device=self.device)
        hdl = symm_mem.rendezvous(t, grp)
        t.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)            # nobody signals before every control region is zero
        self._symm = (t, hdl)
        peers = [t if r == self.rank else hdl.get_buffer(r, (size,), torch.uint8) for r in range(self.world)]
        self._peers = peers
        mine = peers[self.dst]
        for b in range(NBUF):
            whole = mine[b * total:(b + 1) * total]
            self._gathered[b] = whole if self.rank == self.dst else None
            self._local[b] = whole[self.rank * self._cap:(self.rank + 1) * self._cap]
        self._flag_addr = mine.data_ptr() + data + 128 * self.rank          # my completion flag, in dst's memory
        self._flags_local = t.data_ptr() + data                              # dst: flag of rank r at + 128 r
        self._credit_local = t.data_ptr() + data + 128 * self.world          # my credit word, in my memory
        self._credit_table = None
        if self.rank == self.dst:
            addrs = [p.data_ptr() + data + 128 * self.world for p in peers]
            self._credit_table = torch.tensor(addrs, dtype=torch.int64, device=self.device)

    def _views(self, packed, B: int):
        from .solver import BatchedMPC

        return BatchedMPC.unpack(packed, B, self.H, self.fields)

    # typed views are built once per (buffer, shard size): a step must cost the host microseconds, not the ~0.4 ms that
    # 8 ranks x 11 fields of tensor slicing take (measured: that alone capped 8 GPUs at 0.86 ms per step)
    def _local_views(self, buf: int, B: int):
        key = (buf, B)
        v = self._lv_cache.get(key)
        if v is None:
            v = self._lv_cache[key] = self._views(self._local[buf][: packed_layout(B, self.H, self.fields)[1]], B)
        return v

    def _gathered_views(self, buf: int, sizes):
        key = (buf, tuple(sizes))
        out = self._gv_cache.get(key)
        if out is None:
            whole = self._gathered[buf]
            out = {k: [] for k in self.fields}
            for r, (lo, hi) in enumerate(sizes):
                v = self._views(whole[r * self._cap:(r + 1) * self._cap], hi - lo)
                for k in self.fields:
                    out[k].append(v[k])
            self._gv_cache[key] = out
        return out

    # -- device-resident shard ------------------------------------------------------------------------------------
    def submit_device(self, d_paths, d_offsets=None, d_vmax=None, is_localised: bool = False, B_total: Optional[int] = None,
                      warm=None, warm_valid: bool = True) -> _Ticket:
        """Solve this rank's shard (device tensors) and start the exchange; asynchronous.  Every rank calls it."""
        torch, dist = self._torch, self._dist
        B = int(d_paths.shape[0])
        if B_total is None:
            B_total = B * self.world
        sizes = self._size_cache.get(B_total)
        if sizes is None:
            sizes = [shard_range(B_total, r, self.world) for r in range(self.world)]
        lo, hi = sizes[self.rank]
        if hi - lo != B:
            raise ValueError(f"rank {self.rank} holds {B} instances, shard_range says {hi - lo}")
        self._ensure(max(h - l for l, h in sizes))
        self._size_cache[B_total] = sizes
        buf = self._step % NBUF
        main = torch.cuda.current_stream(self.device) if self.cuda else None
        prev2 = self._sig_done[(self._step - 2) % NBUF] if (self._step >= 2 and self.transport != "peer") else None
        if self.cuda and prev2 is not None:
            # buffer `buf` was last used by step - 3; the exchange of step - 2 (which dst joined only after consuming
            # step - 3, and which has overlapped the kernels of step - 1) must be complete before it is overwritten
            main.wait_event(prev2)
        views = self._local_views(buf, B)
        if self.transport == "peer":
            step = self._step
            # my kernels' last CTA raises my flag on dst to step + 1; dst's launch releases every step < `step`; buffer
            # `buf` held step - NBUF, so my first kernel's CTAs start only once my (local) credit word says dst has
            # released it (polled inside the kernel: no stream op between consecutive launches)
            self.solver.attach_completion(self._flag_addr, step + 1,
                                          self._credit_table.data_ptr() if self._credit_table is not None else 0,
                                          self.world if self._credit_table is not None else 0, step,
                                          self._credit_local if step >= NBUF else 0, max(step - NBUF + 1, 0))
            try:
                self.solver.solve_device(d_paths, d_offsets, d_vmax, is_localised, out=views, warm=warm, warm_valid=warm_valid)
            except Exception:
                self.solver.attach_completion()      # nothing was launched: do not leave the one-shot attachment behind
                raise
            self._step += 1
            self._last_B = B
            return _Ticket(buf, sizes, B_total, step)
        self.solver.solve_device(d_paths, d_offsets, d_vmax, is_localised, out=views, warm=warm, warm_valid=warm_valid)
        if self.world > 1:
            if self.cuda:
                solved = torch.cuda.Event()
                solved.record(main)
                self._side.wait_event(solved)
                with torch.cuda.stream(self._side):
                    self._exchange(buf)
                    done = torch.cuda.Event()
                    done.record(self._side)
                self._sig_done[buf] = done
            else:
                self._exchange(buf)
        self._step += 1
        self._last_B = B
        return _Ticket(buf, sizes, B_total, self._step - 1)

    def _exchange(self, buf: int):
        slots = list(self._gathered[buf].view(self.world, -1).unbind(0)) if self.rank == self.dst else None
        self._dist.gather(self._local[buf], slots, dst=self.dst, group=self.group)

    def wait(self, ticket: _Ticket, stream=None):
        """Make `stream` (default: the current stream) wait for the exchange of `ticket`.  dst: dict name -> list of
        per-rank views of the gathered results (instance order); other ranks: None.  The views stay valid until the
        next-but-one submit."""
        torch = self._torch
        if self.transport == "peer":
            if self.rank == self.dst:      # every rank's completion flag for this step, stream-ordered, no SM involved
                for r in range(self.world):
                    self.solver.stream_wait_value32(self._flags_local + 128 * r, ticket.step + 1, stream)
        elif self.cuda and self.world > 1:
            s = torch.cuda.current_stream(self.device) if stream is None else stream
            s.wait_event(self._sig_done[ticket.buf])
        if self.rank != self.dst:
            return None
        return self._gathered_views(ticket.buf, ticket.sizes)

    def last_local_views(self):
        """Typed views of what this rank's kernels wrote in the most recent step (its own slab)."""
        return self._local_views((self._step - 1) % NBUF, self._last_B)

    def drain(self):
        """Block the host until every submitted step and its exchange are complete."""
        if self.cuda:
            self._torch.cuda.synchronize(self.device)

    # -- host global batch (the call a user makes) ------------------------------------------------------------------
    def solve(self, paths, offsets=None, vmax=None, is_localised: bool = False, local: bool = False,
              B_total: Optional[int] = None, deliver: str = "auto") -> Optional[Dict[str, np.ndarray]]:
        """HOST arrays in, HOST arrays out on dst.  `local=False`: every rank passes the same GLOBAL arrays and takes its
        own shard_range slice; `local=True`: each rank passes only its shard (then B_total = sum over ranks is required
        unless all shards are equal).  Pinned staging buffers are kept between calls; synchronous.

        `deliver`: how the results reach dst's HOST memory.
          "shm"      every rank copies its own shard over its OWN PCIe link into one host shared-memory segment that all
                     ranks map and pin (the way the reference itself publishes results, `mp.Array` shared memory,
                     controller.py:274-280): N links in parallel, nothing crosses NVLink, dst returns views of the segment;
          "exchange" the device exchange (peer stores / NCCL gather) lands everything in dst's HBM and dst copies it out
                     over its one PCIe link (79 MB per step at 8 x 4096 instances);
          "auto"     "shm" on CUDA with more than one rank, else "exchange"."""
        torch = self._torch
        self._want_deliver = deliver
        if deliver == "auto":
            deliver = "shm" if (self.cuda and self.world > 1 and not self._shm_failed) else "exchange"
        paths = np.asarray(paths, dtype=np.float64)
        if local:
            B = paths.shape[0]
            if B_total is None:
                B_total = B * self.world
            lo, hi = shard_range(B_total, self.rank, self.world)
            if hi - lo != B:
                raise ValueError("local shard size does not match shard_range(B_total, rank, world)")
            sl = slice(0, B)
        else:
            B_total = paths.shape[0]
            lo, hi = shard_range(B_total, self.rank, self.world)
            B, sl = hi - lo, slice(lo, hi)
        H = self.H
        pin = self.cuda
        if self._host_in is None or self._host_in[0].shape[0] < B:
            mk = lambda *shape: (torch.empty(shape, dtype=torch.float64).pin_memory() if pin
                                 else torch.empty(shape, dtype=torch.float64))
            self._host_in = (mk(B, H, 3), mk(B), mk(B))
            self._dev_in = tuple(torch.empty_like(t, device=self.device) for t in self._host_in)

        def upload(dst_dev, stage, arr):
            """H2D of arr[sl]: straight from the caller's buffer when it is pinned, through the pinned stage otherwise."""
            src = torch.from_numpy(np.ascontiguousarray(np.asarray(arr, dtype=np.float64)[sl]))
            if not (pin and src.is_pinned()):
                stage.copy_(src)
                src = stage
            dst_dev.copy_(src, non_blocking=True)

        dp, do, dv = (t[:B] for t in self._dev_in)
        hp, ho, hv = (t[:B] for t in self._host_in)
        upload(dp, hp, paths)
        if offsets is not None:
            upload(do, ho, offsets)
        if vmax is not None:
            upload(dv, hv, vmax)
        if deliver == "shm":
            try:
                return self._solve_shm(dp, do if offsets is not None else None, dv if vmax is not None else None,
                                       is_localised, B, B_total, lo, hi)
            except ShmUnavailable as e:          # collective: every rank lands here and takes the exchange path instead
                if self._want_deliver != "auto":
                    raise
                self._shm_failed = True
                self.transport_note += f" host delivery falls back to the device exchange: {e}"
        t = self.submit_device(dp, do if offsets is not None else None, dv if vmax is not None else None, is_localised,
                               B_total=B_total)
        views = self.wait(t)
        if self.rank != self.dst:
            self.drain()
            return None
        # D2H of every (rank, field) slab straight into its place in pinned whole-batch arrays (instance order): no host
        # concatenation.  The returned arrays are views of those buffers, valid until the next solve() call.
        spec = _capi.output_spec(H)
        if self._host_out is None or self._host_out[0] != B_total:
            bufs = {}
            for k in self.fields:
                shp, dt = spec[k]
                h = torch.empty((B_total,) + shp, dtype=getattr(torch, dt))
                bufs[k] = h.pin_memory() if pin else h
            self._host_out = (B_total, bufs)
        bufs = self._host_out[1]
        for k in self.fields:
            for r, (l, u) in enumerate(t.sizes):
                bufs[k][l:u].copy_(views[k][r], non_blocking=True)
        self.drain()
        return {k: bufs[k].numpy() for k in self.fields}

    # -- host delivery through one shared-memory segment ------------------------------------------------------------
    def _ensure_shm(self, B_total: int):
        """Collective: one POSIX shared-memory file (two result buffers + a control page), mapped and pinned by every rank."""
        import os
        import uuid

        torch, dist = self._torch, self._dist
        if self._shm is not None and self._shm["B_total"] == B_total:
            return self._shm
        spec = _capi.output_spec(self.H)
        offs, total = {}, 0
        for k in self.fields:
            shp, dt = spec[k]
            nb = B_total * int(np.prod(shp, dtype=np.int64)) * np.dtype(dt).itemsize
            offs[k] = (total, nb)
            total = (total + nb + 4095) // 4096 * 4096
        size = 2 * total + 4096
        name = [None, True]
        if self.rank == self.dst:
            name[0] = uuid.uuid4().hex
            try:      # a container may cap /dev/shm far below what a large batch needs: say so instead of dying on SIGBUS
                st = os.statvfs("/dev/shm")
                name[1] = st.f_bavail * st.f_frsize > size + (64 << 20)
            except OSError:
                name[1] = False
        dist.broadcast_object_list(name, src=self.dst, group=self.group)
        if not name[1]:
            raise ShmUnavailable(f"/dev/shm cannot hold the {size >> 20} MiB result segment")
        path = f"/dev/shm/acmpc_b200_{name[0]}"
        if self.rank == self.dst:
            with open(path, "wb") as f:
                f.truncate(size)
        dist.barrier(group=self.group)
        mm = np.memmap(path, dtype=np.uint8, mode="r+", shape=(size,))
        dist.barrier(group=self.group)
        if self.rank == self.dst:
            os.unlink(path)                          # the mappings keep the segment alive
        # first touch: every rank writes its own slices before anybody pins the segment, so those pages are allocated on
        # the NUMA node its process runs on (the node its GPU's PCIe link usually hangs off) instead of all on dst's
        lo_, hi_ = shard_range(B_total, self.rank, self.world)
        for b in range(2):
            for k in self.fields:
                shp, dt = spec[k]
                o, nb = offs[k]
                per = nb // max(B_total, 1)
                mm[b * total + o + lo_ * per: b * total + o + hi_ * per] = 0
        if self.rank == self.dst:
            mm[2 * total:] = 0
        dist.barrier(group=self.group)
        t = torch.from_numpy(mm)
        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), size, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister of the shared result segment failed ({rc})")
        bufs = []
        for b in range(2):
            tv, nv = {}, {}
            for k in self.fields:
                shp, dt = spec[k]
                o, nb = offs[k]
                seg = t[b * total + o: b * total + o + nb]
                tv[k] = seg.view(getattr(torch, dt)).view((B_total,) + shp)
                nv[k] = mm[b * total + o: b * total + o + nb].view(dt).reshape((B_total,) + shp)
            bufs.append((tv, nv))
        ctrl = mm[2 * total:].view(np.int64)          # [8 r] = steps rank r has delivered, [8 world] = steps dst has released
        _, plain = self.solver.alloc_device_outputs(self._shard_cap(B_total), self.fields)
        self._shm = dict(B_total=B_total, mm=mm, t=t, bufs=bufs, ctrl=ctrl, step=0, plain=plain)
        return self._shm

    def _shard_cap(self, B_total: int) -> int:
        return max(h - l for l, h in (shard_range(B_total, r, self.world) for r in range(self.world)))

    def _solve_shm(self, dp, do, dv, is_localised, B, B_total, lo, hi):
        import time

        torch = self._torch
        sh = self._ensure_shm(B_total)
        step, ctrl = sh["step"], sh["ctrl"]
        buf = step & 1
        if self.rank == self.dst:
            ctrl[8 * self.world] = step               # results of every earlier call are released (valid until the next call)
        else:
            while ctrl[8 * self.world] < step - 1:    # buffer `buf` held step - 2: dst must have started step - 1
                time.sleep(0)
        views = {k: v[:B] for k, v in sh["plain"].items()}
        self.solver.solve_device(dp, do, dv, is_localised, out=views)
        tv, nv = sh["bufs"][buf]
        for k in self.fields:
            tv[k][lo:hi].copy_(views[k], non_blocking=True)     # this rank's shard, over this rank's PCIe link
        torch.cuda.current_stream(self.device).synchronize()
        ctrl[8 * self.rank] = step + 1
        sh["step"] = step + 1
        if self.rank != self.dst:
            return None
        for r in range(self.world):
            while ctrl[8 * r] < step + 1:
                pass
        return nv

    # -- streams of batches, host to host: every rank runs its own copy-in / compute / copy-out pipeline ---------------
    def submit_host(self, paths, offsets=None, vmax=None, is_localised: bool = False, B_total: Optional[int] = None,
                    depth: int = 2) -> int:
        """Asynchronous `solve(local=True, deliver="shm")` for STREAMS of batches: this rank's shard (host arrays, pinned
        for true asynchrony) goes up on a copy stream, is solved on a compute stream and its results go down this rank's
        own PCIe link into the shared host segment on a third stream; consecutive steps overlap.  Every rank calls it;
        `wait_host(ticket)` on dst returns whole-batch numpy views, valid until `depth` further submits."""
        import time

        torch = self._torch
        paths = np.asarray(paths, dtype=np.float64)
        B = paths.shape[0]
        if B_total is None:
            B_total = B * self.world
        lo, hi = shard_range(B_total, self.rank, self.world)
        if hi - lo != B:
            raise ValueError("local shard size does not match shard_range(B_total, rank, world)")
        sh = self._ensure_shm(B_total)
        if "pipe" not in sh:
            dev = self.device
            slots = []
            for _ in range(depth):
                _, views = self.solver.alloc_device_outputs(B, self.fields)
                slots.append(dict(d_paths=torch.empty((B, self.H, 3), dtype=torch.float64, device=dev),
                                  d_off=torch.empty(B, dtype=torch.float64, device=dev),
                                  d_vmax=torch.empty(B, dtype=torch.float64, device=dev),
                                  h_paths=torch.empty((B, self.H, 3), dtype=torch.float64).pin_memory(),
                                  h_off=torch.empty(B, dtype=torch.float64).pin_memory(),
                                  h_vmax=torch.empty(B, dtype=torch.float64).pin_memory(),
                                  views=views, in_ready=torch.cuda.Event(), solved=torch.cuda.Event(),
                                  out_done=torch.cuda.Event(), used=False))
            sh["pipe"] = dict(slots=slots, depth=depth, streams=[torch.cuda.Stream(device=dev) for _ in range(3)],
                              count=torch.zeros(1, dtype=torch.int64, device=dev),
                              ctrl_t=sh["t"][2 * ((sh["t"].numel() - 4096) // 2):].view(torch.int64))
            if depth != 2:
                raise ValueError("the shared segment holds two result buffers: depth must be 2")
        pp = sh["pipe"]
        step, ctrl = sh["step"], sh["ctrl"]
        s = pp["slots"][step % pp["depth"]]
        s_in, s_run, s_out = pp["streams"]
        if self.rank == self.dst:
            ctrl[8 * self.world] = step - pp["depth"] + 1      # steps <= step - depth are released
        else:
            while ctrl[8 * self.world] < step - pp["depth"] + 1:
                time.sleep(0)

        def stage(buf, arr):
            src = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
            if src.is_pinned():
                return src
            if s["used"]:
                s["in_ready"].synchronize()     # the previous upload from this staging buffer has left it
            buf.copy_(src)
            return buf

        with torch.cuda.stream(s_in):
            if s["used"]:
                s_in.wait_event(s["solved"])
            s["d_paths"].copy_(stage(s["h_paths"], paths), non_blocking=True)
            if offsets is not None:
                s["d_off"].copy_(stage(s["h_off"], offsets), non_blocking=True)
            if vmax is not None:
                s["d_vmax"].copy_(stage(s["h_vmax"], vmax), non_blocking=True)
            s["in_ready"].record(s_in)
        with torch.cuda.stream(s_run):
            s_run.wait_event(s["in_ready"])
            if s["used"]:
                s_run.wait_event(s["out_done"])
            self.solver.solve_device(s["d_paths"], s["d_off"] if offsets is not None else None,
                                     s["d_vmax"] if vmax is not None else None, is_localised, out=s["views"], stream=s_run)
            s["solved"].record(s_run)
        key = ("dst", step % 2, step % pp["depth"])
        pairs = pp.get(key)
        if pairs is None:      # (target slice in the segment, device source) per field, built once per buffer pair
            tv, _ = sh["bufs"][step % 2]
            pairs = pp[key] = [(tv[k][lo:hi], s["views"][k]) for k in self.fields]
            pp[("cnt", step % 2)] = pp["ctrl_t"][8 * self.rank: 8 * self.rank + 1]
        with torch.cuda.stream(s_out):
            s_out.wait_event(s["solved"])
            for dst_t, src_t in pairs:
                dst_t.copy_(src_t, non_blocking=True)
            # the "delivered" counter travels down the same stream AFTER the data, so it lands after them
            pp["count"].fill_(step + 1)
            pp["ctrl_t"][8 * self.rank: 8 * self.rank + 1].copy_(pp["count"], non_blocking=True)
            s["out_done"].record(s_out)
        s["used"] = True
        sh["step"] = step + 1
        return step

    def wait_host(self, ticket: int):
        """dst: block until every rank has delivered step `ticket`, return the whole-batch numpy views; others: None."""
        sh = self._shm
        if self.rank != self.dst:
            return None
        ctrl = sh["ctrl"]
        for r in range(self.world):
            while ctrl[8 * r] < ticket + 1:
                pass
        return sh["bufs"][ticket % 2][1]
