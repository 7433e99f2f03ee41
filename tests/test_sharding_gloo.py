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
