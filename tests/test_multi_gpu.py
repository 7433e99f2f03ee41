"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): the product-level sharded call ShardedMPC on NCCL, one process per
GPU -- both transports of the final exchange ("peer": the kernels store straight into rank 0's slab over NVLink;
"nccl": one gather) must deliver on rank 0 exactly what one GPU computes for the whole batch, bit for bit, over more
steps than there are buffers, for even and ragged shards, from host arrays and from device-resident shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

FIELDS = ["controls", "prediction", "cum_time", "status", "status_speed", "iters", "cost", "derived"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, transport, B, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from ac_mpc_b200 import BatchedMPC, _capi, sharding, tracks
        from ac_mpc_b200.sharded import ShardedMPC

        cfg = _capi.default_config()
        paths, vmax = tracks.perturbed_batch("monza", B, seed=8)
        offs = np.linspace(-0.4, 0.4, B)
        sh = ShardedMPC(cfg, fields=FIELDS, device=rank, transport=transport)
        ok, msgs = True, []
        want = BatchedMPC(cfg, device=rank).solve_host(paths, offs, vmax, False, fields=FIELDS) if rank == 0 else None
        for rep in range(6):
            # host delivery alternates between the shared-memory segment (every rank over its own PCIe link) and the
            # device exchange + one D2H on rank 0
            out = sh.solve(paths, offs, vmax, False, deliver="shm" if rep & 1 else "exchange")
            if rank == 0:
                for k in FIELDS:
                    if not np.array_equal(out[k], want[k]):
                        ok = False
                        msgs.append(f"rep {rep} field {k}")
        lo_, hi_ = sharding.shard_range(B, rank, world)
        for rep in range(3):                       # shm delivery back to back (both buffers, release counter), local shards
            out = sh.solve(paths[lo_:hi_], offs[lo_:hi_], vmax[lo_:hi_], False, local=True, B_total=B)
            if rank == 0 and not all(np.array_equal(out[k], want[k]) for k in FIELDS):
                ok = False
                msgs.append(f"shm local rep {rep}")
        # streams of batches host to host: every rank its own 3-stream pipeline, results through the shared segment
        prev, n_ok = None, 0
        for rep in range(6):
            sc = 1.0 + 0.001 * rep                 # a different batch every step (scaled lateral offsets)
            tk = sh.submit_host(paths[lo_:hi_], offs[lo_:hi_] * sc, vmax[lo_:hi_], False, B_total=B)
            if prev is not None:
                res = sh.wait_host(prev[0])
                if rank == 0:
                    ref = BatchedMPC(cfg, device=rank).solve_host(paths, offs * prev[1], vmax, False, fields=["controls", "iters"])
                    n_ok += int(np.array_equal(res["controls"], ref["controls"]) and np.array_equal(res["iters"], ref["iters"]))
            prev = (tk, sc)
        res = sh.wait_host(prev[0])
        torch.cuda.synchronize()
        if rank == 0 and n_ok != 5:
            ok = False
            msgs.append(f"submit_host / wait_host: {n_ok} of 5 pipelined batches equal the single-GPU solve")
        # device-resident shards, asynchronous, several steps in flight; only the last result is read
        lo, hi = sharding.shard_range(B, rank, world)
        dp = torch.from_numpy(paths[lo:hi]).cuda()
        do = torch.from_numpy(offs[lo:hi]).cuda()
        dv = torch.from_numpy(vmax[lo:hi]).cuda()
        for rep in range(7):
            t = sh.submit_device(dp, do, dv, False, B_total=B)
        views = sh.wait(t)
        sh.drain()
        if rank == 0:
            for k in FIELDS:
                got = torch.cat(views[k], dim=0).cpu().numpy()
                if not np.array_equal(got, want[k]):
                    ok = False
                    msgs.append(f"device path field {k}")
        q.put((rank, ok, sh.transport, sh.transport_note, msgs))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport,B", [("nccl", 4096), ("nccl", 1001), ("peer", 4096), ("peer", 1001), ("auto", 2048)])
def test_sharded_solve_on_two_gpus_equals_one_gpu(transport, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, transport, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, ok, used, note, msgs in got:
        assert ok, f"rank {rank} ({used}): {msgs}"
        if transport != "auto":
            assert used == transport, note
