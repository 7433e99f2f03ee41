#!/usr/bin/env python
"""What does the "peer" transport cost a kernel?  2 ranks; each times its own 4096-instance step (library events) with the
outputs in (a) plain device memory, (b) its OWN symmetric-memory buffer, (c) the PEER's symmetric-memory buffer, each without
and with the completion flag / credit protocol.  Experiment tooling (torchrun --nproc-per-node 2 tools/peer_cost_probe.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ac_mpc_b200 import BatchedMPC, _capi, tracks  # noqa: E402
from ac_mpc_b200.sharding import packed_layout  # noqa: E402

FIELDS = ["controls", "prediction", "cum_time", "v_ref", "cost", "pri_res", "dua_res", "status", "status_speed", "iters",
          "rho_updates"]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

B, H = 4096, 50
paths, vmax = tracks.perturbed_batch("monza", B, seed=1 + rank)
mpc = BatchedMPC(_capi.default_config(), device=rank)
dp, dv = torch.from_numpy(paths).to(dev), torch.from_numpy(vmax).to(dev)
cap = packed_layout(B, H, FIELDS)[1]
plain = torch.empty(cap, dtype=torch.uint8, device=dev)
t = symm_mem.empty(cap + 1024, dtype=torch.uint8, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
t.zero_()
torch.cuda.synchronize()
dist.barrier()
peer = hdl.get_buffer(1 - rank, (cap + 1024,), torch.uint8)
table = torch.tensor([t.data_ptr() + cap + 512, peer.data_ptr() + cap + 512], dtype=torch.int64, device=dev)


def run(buf, flags, label):
    views = BatchedMPC.unpack(buf[:cap], B, H, FIELDS)
    flag_addr = buf.data_ptr() + cap + 128 * rank
    mpc.set_profiling(True)
    step = [0]
    for it in range(12):
        if flags:
            step[0] += 1
            mpc.attach_completion(flag_addr, step[0], table.data_ptr(), 2, step[0], t.data_ptr() + cap + 512, 0)
        mpc.solve_device(dp, None, dv, False, out=views)
        if it == 1:
            torch.cuda.synchronize()
            mpc.collect_kernel_ms()
    torch.cuda.synchronize()
    k = mpc.collect_kernel_ms()
    mpc.set_profiling(False)
    dist.barrier()
    n = k["launches"]
    print(f"rank {rank} {label:44s} speed {k['speed_ms'] / n:.4f} control {k['control_ms'] / n:.4f} ms", flush=True)


for buf, name in ((plain, "plain cudaMalloc memory"), (t, "own symmetric-memory buffer"), (peer, "PEER's symmetric-memory buffer (NVLink)")):
    if buf is plain:
        run(buf, False, name)
        continue
    run(buf, False, name + ", no flags")
    run(buf, True, name + ", flags+credits")
dist.destroy_process_group()
