"""BASELINE config 5 check under torchrun (one process per GPU): the tile-sharded global grid against the whole map computed locally.
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/tiled_elevation_check.py"""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import torch.distributed as dist

from mr_slam_b200 import _lib, elevation, synth
from mr_slam_b200 import dist as mdist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_lib.init(local)
L, res = 1200, 0.1
sync = lambda: (_lib.lib().mrslam_device_synchronize(), torch.cuda.synchronize())
whole = elevation.ElevationMap()
whole.Init_GPU_elevationmap(L, res, 2.5, 0.2)
m = elevation.ElevationMap()
m.Init_GPU_elevationmap(L, res, 2.5, 0.2)
tiled = mdist.TiledElevationMap(m, L, device="cuda", sync=sync)
frames = []
for f in range(4):
    pose = synth.pose_matrix(2.0 * f + 0.3, -1.0 * f, 0.1 * f, 0.3 * f)
    pts = synth.make_scan(pose, "hdl64", scan_id=100 + f)
    ei = synth.elevation_inputs(pose)
    args = (ei["T"], -5.0, 5.0, ei["min_r"], ei["beam_a"], ei["beam_c"], ei["sensorJacobian"], ei["rotationVariance"],
            ei["C_SB_transpose"], ei["P_mul_C_BM_transpose"], ei["B_r_BS_skew"])
    d = [torch.from_numpy(np.ascontiguousarray(pts[:, i])).cuda() for i in range(3)]
    frames.append((pose[:3, 3].astype(np.float32), d, len(pts), args))
torch.cuda.synchronize()
for mm in (whole, m):
    for pos, d, n, args in frames:
        mm.Move(pos, res, L)
        mm.process_fuse_device(d[0], d[1], d[2], n, *args, near_filter=True, want_count=False)
t0 = time.time()
got = tiled.gather_full(("elevation", "variance"))
feat = tiled.map_feature_owned()
dt = time.time() - t0
W = whole.get_map()
FW = whole.Map_feature(L)
ok = np.array_equal(got["elevation"].cpu().numpy(), W["elevation"]) and np.array_equal(got["variance"].cpu().numpy(), W["variance"])
for k in ("rough", "slope", "traver"):
    ok = ok and np.array_equal(feat[k], FW[k].reshape(L, L)[tiled.row_lo:tiled.row_hi])
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("tiled global elevation (world %d, rows per tile %s): union == whole map and owned-row features == whole-map features: %s "
          "(gather + halo + features %.1f ms)" % (world, [b - a for a, b in zip(tiled.bounds[:-1], tiled.bounds[1:])],
                                                  "OK" if t.item() == 1.0 else "MISMATCH", 1e3 * dt))
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
