"""Device / call times of the SURVEY 8f ("next") rows at bench sizes -> one JSON line (copied to profiles/ by hand):
keyframe merge (6 x 120k-pt keyframes, 0.2 m voxels), solve_translation, calcRelOri vs 256 signatures, RING++ front end
(kNN(30) + eigen features of a 30k-pt cloud), PointCloud2 unpack (120k records), Map_feature and Raytracing on a 1200^2 map.
Wall times of the python-level call (inputs on the device where the API allows it), median of 10 after 3 warm-ups."""
import json, os, sys, time
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
from mr_slam_b200 import _lib, elevation, msgs, pygicp, ring, synth

_lib.init(0)
dev = torch.device("cuda", 0)
L = _lib.lib()


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    L.mrslam_device_synchronize(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        L.mrslam_device_synchronize(); torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts))


src, tgt, *_ = synth.make_pair(0, "hdl64")
out = {}
kfs = [torch.from_numpy(np.concatenate([src + np.float32(0.05 * i), np.ones((len(src), 1), np.float32)], 1)).to(dev) for i in range(6)]
Ts = np.stack([synth.pose_matrix(0.3 * i, 0.1 * i, 0, 0.02 * i) for i in range(6)]).astype(np.float32)
out["keyframe_merge_6x120k_ms"] = timed(lambda: pygicp.merge_nearest_keyframes(kfs, Ts, as_cloud=True))
scans = torch.from_numpy(np.stack([src + np.array([0, 0, 1.8], np.float32)] * 2)).to(dev)
bev = ring.scans_to_bev_device(scans)
sino, tir = ring.generate_RING_batch(bev, want_sino=True)
out["solve_translation_ms"] = timed(lambda: ring.solve_translation(sino[0], sino[1], 3))
rng = np.random.default_rng(0)
new = tuple(torch.from_numpy(rng.standard_normal((40, 120)).astype(np.float32)).to(dev) for _ in range(2))
old = tuple(torch.from_numpy(rng.standard_normal((256, 40, 120)).astype(np.float32)).to(dev) for _ in range(2))
out["calc_rel_ori_256_signatures_ms"] = timed(lambda: ring.calc_rel_ori(new, old))
pc30 = src[::4].copy()
out["ringpp_build_neighbors_30k_k30_ms"] = timed(lambda: ring.build_neighbors_NN(pc30, 30), reps=5)
rec = np.zeros((len(src), 8), np.float32); rec[:, :3] = src; rec[:, 4] = 1.0
blob = torch.from_numpy(rec.view(np.uint8).reshape(-1)).to(dev)
out["pointcloud2_unpack_120k_ms"] = timed(lambda: msgs.pointcloud2_unpack(blob, len(src), 32, dict(x=0, y=4, z=8, intensity=16)))
em = elevation.ElevationMap(); em.Init_GPU_elevationmap(1200, 0.1, 2.5, 0.2)
frames, keep = [], []
for f in range(8):
    pose = synth.pose_matrix(0.4 * f, -0.2 * f, 0.0, 0.05 * f)
    e = synth.elevation_inputs(pose)
    d = [torch.from_numpy(np.ascontiguousarray(src[:, i])).to(dev) for i in range(3)]
    keep.append(d)
    frames.append(elevation.ElevationMap.make_frame(d[0], d[1], d[2], len(src), e["T"], -5.0, 5.0, e["min_r"], e["beam_a"], e["beam_c"], e["sensorJacobian"],
                                                    e["rotationVariance"], e["C_SB_transpose"], e["P_mul_C_BM_transpose"], e["B_r_BS_skew"],
                                                    near_filter=True, position=pose[:3, 3].astype(np.float32)))
em.process_fuse_device_batch(frames, 0.1, 1200)
out["map_feature_1200sq_incl_9_plane_readout_ms"] = timed(lambda: em.Map_feature(1200), reps=5)
def ray():
    em.Raytracing(1200)
out["raytracing_1200sq_ms"] = timed(ray, reps=5)
print(json.dumps(out))
