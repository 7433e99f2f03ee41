"""CPU, world size 2, gloo: the N > 1 host logic -- contiguous sharding of a batch over ranks and the
single final gather of the packed result buffers.  The per-rank 'solve' is faked
by a deterministic fill (no compute without a GPU); what is under test is that every instance's slab
lands at the right place, for even and ragged batch sizes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ac_mpc_b200 import _capi, sharding

H = 12
FIELDS = ["controls", "cum_time", "cost", "status", "iters"]


def _fake_result(field, lo, hi):
    shp, dt = _capi.output_spec(H)[field]
    per = int(np.prod(shp, dtype=np.int64))
    base = np.arange(lo, hi, dtype=np.float64)[:, None] * 1000.0 + np.arange(per)[None, :] + len(field)
    return base.reshape((hi - lo,) + shp).astype(dt)


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(B, rank, world)
        offs, total = sharding.packed_layout(hi - lo, H, FIELDS)
        packed = torch.zeros(total, dtype=torch.uint8)
        for k in FIELDS:
            o, nb = offs[k]
            packed[o:o + nb] = torch.from_numpy(_fake_result(k, lo, hi).reshape(-1).view(np.uint8).copy())
        views = sharding.gather_packed(packed, B, H, FIELDS)
        whole = sharding.concat_views(views)
        ok = all(np.array_equal(whole[k], _fake_result(k, 0, B)) for k in FIELDS)
        q.put((rank, ok, {k: whole[k].shape for k in FIELDS}))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_the_batch():
    for B in (0, 1, 7, 64, 4096, 4099):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


@pytest.mark.parametrize("B", [8, 9])
def test_world2_gather_of_packed_results(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shapes in got:
        assert ok, f"rank {rank}: gathered batch differs"
        assert shapes["controls"] == (B, 2, H - 1)


# ---- the product-level call: ShardedMPC (ac_mpc_b200/sharded.py) over gloo with the ORACLE standing in for the GPU ----
class _OracleSolver:
    """Stand-in for BatchedMPC on a GPU-less box (tests only): same solve_device / unpack surface, results from the CPU
    oracle.  `device = None` tells ShardedMPC to run its exchange on CPU tensors (gloo)."""

    device = None

    def __init__(self, cfg_kw):
        from oracle import port

        self._port, self._cfg = port, port.default_config(**cfg_kw)

    def solve_device(self, paths, offsets=None, vmax=None, is_localised=False, out=None, warm=None, warm_valid=True):
        res = self._port.solve_batch(self._cfg, paths.numpy(), None if offsets is None else offsets.numpy(),
                                     None if vmax is None else vmax.numpy(), is_localised, nthreads=2)
        for k, v in out.items():
            v.copy_(torch.from_numpy(res[k]))
        return out


def _sharded_worker(rank, world, port_no, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ac_mpc_b200 import tracks
        from ac_mpc_b200.sharded import ShardedMPC
        from oracle import port

        kw = dict(horizon=20)
        fields = ["controls", "status", "iters", "cost"]
        paths, vmax = tracks.perturbed_batch("monza", B, horizon=20, seed=6)
        offs = np.linspace(-0.3, 0.3, B)
        sh = ShardedMPC(_capi.default_config(**kw), fields=fields, solver=_OracleSolver(kw))
        ok = True
        for rep in range(4):                       # more steps than buffers: the rotation is exercised
            out = sh.solve(paths, offs, vmax, bool(rep & 1))
            if rank == 0:
                want = port.solve_batch(port.default_config(**kw), paths, offs, vmax, bool(rep & 1), nthreads=2)
                ok = ok and all(np.array_equal(out[k], want[k]) for k in fields) and out["controls"].shape == (B, 2, 19)
            else:
                ok = ok and out is None
        # only this rank's shard passed in (local=True)
        lo, hi = sharding.shard_range(B, rank, world)
        out = sh.solve(paths[lo:hi], offs[lo:hi], vmax[lo:hi], False, local=True, B_total=B)
        if rank == 0:
            want = port.solve_batch(port.default_config(**kw), paths, offs, vmax, False, nthreads=2)
            ok = ok and all(np.array_equal(out[k], want[k]) for k in fields)
        q.put((rank, bool(ok), sh.transport))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [16, 21])
def test_world2_sharded_solve_returns_the_whole_batch_on_rank0(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port_no, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, transport in got:
        assert ok, f"rank {rank}: sharded result differs from the unsharded oracle batch"
        assert transport == "nccl"          # the collective transport (gloo here); "peer" needs NVLink
