import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from ac_mpc_b200 import tracks
from ac_mpc_b200.control import build_mpc
import bench
B,H=4096,50
paths,vmax=bench.workload("monza",B,H,0)
veh = type("V", (), {"vehicle_data": type("D", (), {"wheelbase": 2.65, "width": 1.99})(), "max_steering_angle": lambda self: 0.30})()
api = build_mpc(tracks.racing_config("monza", H), veh, device=0)
hp=torch.from_numpy(paths).pin_memory(); hv=torch.from_numpy(vmax).pin_memory()
out=api._batched().alloc_host_outputs(B, bench.BENCH_FIELDS, pinned=True)
for _ in range(5): api.get_control_batch(hp.numpy(), None, hv.numpy(), False, out=out)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(100): api.get_control_batch(hp.numpy(), None, hv.numpy(), False, out=out)
dt=time.perf_counter()-t0
print("e2e %.0f solves/s  %.4f ms/step"%(B*100/dt, dt*10))
