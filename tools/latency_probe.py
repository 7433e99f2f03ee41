"""Batch-1 breakdown: device time of the two kernels (library events) next to the host-side call time."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from ac_mpc_b200 import BatchedMPC, _capi, tracks
cl = tracks.synthetic_centreline("monza")
p1 = tracks.make_instances(cl, [0], 50)
mpc = BatchedMPC(_capi.default_config(), device=0)
d = torch.from_numpy(p1).cuda()
packed, views = mpc.alloc_device_outputs(1)
for _ in range(10): mpc.solve_device(d, out=views)
torch.cuda.synchronize()
mpc.set_profiling(True)
for _ in range(50): mpc.solve_device(d, out=views)
torch.cuda.synchronize()
k = mpc.collect_kernel_ms()
print("device: speed %.1f us, control %.1f us per call; iters" % (k["speed_ms"] / k["launches"] * 1e3, k["control_ms"] / k["launches"] * 1e3), views["iters"].cpu().numpy())
mpc.set_profiling(False)
o = mpc.alloc_host_outputs(1, pinned=True)
ts = []
for _ in range(200):
    t0 = time.perf_counter(); mpc.solve_host(p1, out=o); ts.append(time.perf_counter() - t0)
print("host call p50 %.1f us" % (np.median(ts[20:]) * 1e6))
o2 = mpc.alloc_host_outputs(1, ["controls", "prediction", "cum_time", "status"], pinned=True)
ts = []
for _ in range(200):
    t0 = time.perf_counter(); mpc.solve_host(p1, out=o2); ts.append(time.perf_counter() - t0)
print("host call, 4 output fields, p50 %.1f us" % (np.median(ts[20:]) * 1e6))
